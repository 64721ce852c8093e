// api.cu -- context life cycle, error reporting, pinned memory and the host-side hash helpers of the C ABI
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "fc_internal.cuh"

void fc_genome_release(fc_ctx* ctx);
void fc_agg_release(fc_ctx* ctx);

static thread_local std::string g_create_error;

int fc_fail(fc_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_create_error = buf;
  return code;
}

extern "C" int fc_abi_version(void) { return FC_ABI_VERSION; }

extern "C" int fc_ctx_create(int device, fc_ctx** out) {
  if (!out) return FC_E_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fc_fail(nullptr, FC_E_CUDA, "no CUDA device available (%s); libfindcirc_b200 has no CPU path",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count) return fc_fail(nullptr, FC_E_ARG, "device %d out of range (0..%d)", device, count - 1);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fc_fail(nullptr, FC_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  fc_ctx* ctx = new fc_ctx();
  ctx->device = device;
  e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return fc_fail(nullptr, FC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  if (const char* e = getenv("FC_HOST_CHUNK")) {
    const long long v = atoll(e);
    if (v >= 1024) ctx->host_chunk = v;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (ctx->sm_count <= 0) ctx->sm_count = 148;
  *out = ctx;
  return FC_OK;
}

extern "C" void fc_ctx_destroy(fc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  fc_genome_release(ctx);
  fc_agg_release(ctx);
  for (auto& b : ctx->host_path) b.release();
  ctx->tie_off.release();
  for (auto& b : ctx->pk) b.release();
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->own_stream2) {
    cudaStreamDestroy(ctx->own_stream2);
    cudaEventDestroy(ctx->ev_chunk[0]);
    cudaEventDestroy(ctx->ev_chunk[1]);
  }
  delete ctx;
}

extern "C" const char* fc_last_error(fc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" void* fc_pinned_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0) return nullptr;
  if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void fc_pinned_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int fc_device_sync(fc_ctx* ctx) {
  if (!ctx) return FC_E_ARG;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  FC_CUDA(ctx, cudaDeviceSynchronize());
  return FC_OK;
}

extern "C" int64_t fc_launch_count(fc_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------ hashing (host)
static inline uint64_t fnv1a(const uint8_t* p, int64_t n, uint64_t h) {
  for (int64_t i = 0; i < n; ++i) {
    h ^= p[i];
    h *= 0x100000001b3ULL;
  }
  return h;
}

extern "C" uint64_t fc_hash_bytes(const uint8_t* p, int64_t n) { return fc_mix64(fnv1a(p, n, 0xcbf29ce484222325ULL) + (uint64_t)n); }

// The reference keeps {(read, sample), (rev_comp(read), sample)} per junction and reports len/2 (find_circ.py:581-590):
// a strand-invariant key is min(hash(read), hash(revcomp(read))); bit 0 of the result flags reads that equal their own
// reverse complement (they contribute ONE set element, not two).
extern "C" uint64_t fc_hash_read(const uint8_t* seq, int64_t n, int32_t* is_palindrome) {
  static uint8_t comp[256];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 256; ++i) comp[i] = (uint8_t)i;
    const char* a = "ACGTKMRYSWBVHDNacgtkmryswbvhdn";
    const char* b = "TGCAMKYRSWVBDHNtgcamkyrswvbdhn";
    for (int i = 0; a[i]; ++i) comp[(uint8_t)a[i]] = (uint8_t)b[i];
    init = true;
  }
  uint64_t hf = 0xcbf29ce484222325ULL, hr = 0xcbf29ce484222325ULL;
  bool pal = true;
  for (int64_t i = 0; i < n; ++i) {
    uint8_t f = seq[i], r = comp[seq[n - 1 - i]];
    hf = (hf ^ f) * 0x100000001b3ULL;
    hr = (hr ^ r) * 0x100000001b3ULL;
    pal = pal && (f == r);
  }
  hf = fc_mix64(hf + (uint64_t)n);
  hr = fc_mix64(hr + (uint64_t)n);
  uint64_t h = hf < hr ? hf : hr;
  h = (h & ~1ull) | (pal ? 1ull : 0ull);
  if (is_palindrome) *is_palindrome = pal ? 1 : 0;
  return h;
}

// the same on the device: one thread per read (rows of a fixed-stride byte matrix in device memory)
__constant__ uint8_t c_comp[256];
__global__ void hash_reads_kernel(int64_t n, const uint8_t* __restrict__ seq, int32_t stride, const int32_t* __restrict__ len,
                                  int32_t fixed_len, uint64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* row = seq + i * (int64_t)stride;
  const int m = len ? len[i] : fixed_len;
  uint64_t hf = 0xcbf29ce484222325ULL, hr = 0xcbf29ce484222325ULL;
  bool pal = true;
  for (int k = 0; k < m; ++k) {
    const uint8_t f = row[k], r = c_comp[row[m - 1 - k]];
    hf = (hf ^ f) * 0x100000001b3ULL;
    hr = (hr ^ r) * 0x100000001b3ULL;
    pal = pal && (f == r);
  }
  hf = fc_mix64(hf + (uint64_t)m);
  hr = fc_mix64(hr + (uint64_t)m);
  uint64_t h = hf < hr ? hf : hr;
  out[i] = (h & ~1ull) | (pal ? 1ull : 0ull);
}

extern "C" int fc_hash_reads_device(fc_ctx* ctx, int64_t n, const uint8_t* d_seq, int32_t stride, const int32_t* d_len,
                                    int32_t fixed_len, uint64_t* d_out, void* stream) {
  if (!ctx || n < 0 || !d_seq || !d_out || (!d_len && fixed_len < 0)) return FC_E_ARG;
  if (n == 0) return FC_OK;
  static bool table_up = false;  // (per process and device image; the table is constant)
  if (!table_up) {
    uint8_t comp[256];
    for (int i = 0; i < 256; ++i) comp[i] = (uint8_t)i;
    const char* a = "ACGTKMRYSWBVHDNacgtkmryswbvhdn";
    const char* b = "TGCAMKYRSWVBDHNtgcamkyrswvbdhn";
    for (int i = 0; a[i]; ++i) comp[(uint8_t)a[i]] = (uint8_t)b[i];
    FC_CUDA(ctx, cudaMemcpyToSymbol(c_comp, comp, 256));
    table_up = true;
  }
  hash_reads_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, d_seq, stride, d_len, fixed_len, d_out);
  FC_LAUNCH_CHECK(ctx);
  return FC_OK;
}

extern "C" int fc_hash_reads_host(int64_t n, const uint8_t* h_seq, int32_t stride, const int32_t* h_len, uint64_t* h_out,
                                  uint8_t* h_pal) {
  if (n < 0 || !h_seq || !h_len || !h_out) return FC_E_ARG;
  for (int64_t i = 0; i < n; ++i) {
    int32_t pal = 0;
    h_out[i] = fc_hash_read(h_seq + i * (int64_t)stride, h_len[i], &pal);
    if (h_pal) h_pal[i] = (uint8_t)pal;
  }
  return FC_OK;
}
