// genome.cu -- device-resident genome store.
//
// Replaces Track / GenomeAccessor / indexed_fasta (/root/reference/find_circ.py:103-215, 242-371): instead of an
// mmap'ed FASTA sliced per request, every chromosome is packed once (on the GPU) into three bit planes
//   plo / phi   low / high bit of the 2-bit base code (A0 C1 G2 T3)
//   pn          1 = not ACGT -> reads as 'N'
// laid out in ONE global coordinate space with FC_GENOME_PAD bases of N between chromosomes, so that windows
// hanging over a chromosome end read 'N' exactly as find_circ.py:194-211 pads them.
// For the scan the planes are additionally re-cut into overlapping sector-sized tiles (see scan_core.cuh):
// 180 GB of HBM make the 2-4x replication of a mammalian genome a non-issue, and it turns every window fetch
// into a single 256-bit load per 32-byte sector with the N summary in-band.
#include <stdarg.h>
#include <string.h>

#include <algorithm>

#include "fc_internal.cuh"

namespace {

constexpr int64_t ALIGN = 2048;          // chromosome offsets
constexpr int64_t CHUNK = 32ll << 20;    // ASCII staging chunk (bases), multiple of ALIGN
constexpr int64_t SLACK_BASES = 4096;    // readable bases past `total` (tile building and unaligned window loads)

// one thread per 32-base word of the chunk: 32 ASCII bytes in, one word per plane out
__global__ void pack_chunk_kernel(const uint8_t* __restrict__ ascii, int64_t chunk_bases /*valid bytes in ascii*/,
                                  int64_t gbase /*global base index of ascii[0], multiple of 32*/, int64_t n_words,
                                  uint32_t* __restrict__ plo, uint32_t* __restrict__ phi, uint32_t* __restrict__ pn,
                                  unsigned long long* __restrict__ counters) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned n_n = 0, n_other = 0;
  if (w < n_words) {
    int64_t s = w * 32;
    uint32_t lo = 0, hi = 0, nn = 0;
    if (s + 32 <= chunk_bases) {
      const uint4* src = reinterpret_cast<const uint4*>(ascii + s);
      uint4 v[2] = {src[0], src[1]};
      const uint8_t* b = reinterpret_cast<const uint8_t*>(v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        uint32_t c = b[j] & 0xDFu;
        uint32_t code = 0, isn = 1;
        if (c == 'A') { code = 0; isn = 0; }
        else if (c == 'C') { code = 1; isn = 0; }
        else if (c == 'G') { code = 2; isn = 0; }
        else if (c == 'T') { code = 3; isn = 0; }
        else if (c == 'N') { n_n++; }
        else { n_other++; }
        lo |= (code & 1u) << j;
        hi |= (code >> 1) << j;
        nn |= isn << j;
      }
    } else {
      for (int j = 0; j < 32; ++j) {
        int64_t p = s + j;
        uint32_t code = 0, isn = 1;
        if (p < chunk_bases) {
          uint32_t c = ascii[p] & 0xDFu;
          if (c == 'A') { code = 0; isn = 0; }
          else if (c == 'C') { code = 1; isn = 0; }
          else if (c == 'G') { code = 2; isn = 0; }
          else if (c == 'T') { code = 3; isn = 0; }
          else if (c == 'N') { n_n++; }
          else { n_other++; }
        }
        lo |= (code & 1u) << j;
        hi |= (code >> 1) << j;
        nn |= isn << j;
      }
    }
    int64_t gw = (gbase >> 5) + w;
    plo[gw] = lo;
    phi[gw] = hi;
    pn[gw] = nn;
  }
  for (int o = 16; o; o >>= 1) {
    n_n += __shfl_down_sync(0xffffffffu, n_n, o);
    n_other += __shfl_down_sync(0xffffffffu, n_other, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_n) atomicAdd(&counters[0], (unsigned long long)n_n);
    if (n_other) atomicAdd(&counters[1], (unsigned long long)n_other);
  }
}

__device__ inline uint32_t extract32(const uint32_t* __restrict__ plane, int64_t bit) {
  const uint32_t* p = plane + (bit >> 5);
  return __funnelshift_r(p[0], p[1], (uint32_t)(bit & 31));
}

// one thread per tile: cut 4T words per plane out of the master planes at base t*S, put the N summary in the
// 4 spare bits of the lo plane's last word
__global__ void build_tiles_kernel(const uint32_t* __restrict__ plo, const uint32_t* __restrict__ phi,
                                   const uint32_t* __restrict__ pn, int64_t n_tiles, int T, int S,
                                   uint32_t* __restrict__ tiles) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int PW = 4 * T;
  const int64_t g0 = t * (int64_t)S;
  uint32_t* dst = tiles + t * (int64_t)(8 * T);
  uint32_t anyn = 0;
  for (int j = 0; j < PW; ++j) {
    uint32_t lo = extract32(plo, g0 + 32 * j);
    uint32_t hi = extract32(phi, g0 + 32 * j);
    uint32_t nn = extract32(pn, g0 + 32 * j);
    if (j == PW - 1) {
      lo &= 0x0FFFFFFFu;
      hi &= 0x0FFFFFFFu;
      nn &= 0x0FFFFFFFu;
    }
    anyn |= nn;
    if (j == PW - 1 && anyn) lo |= fc::TILE_FLAG_N;
    if (j < PW - 1) {
      dst[j] = lo;
    } else {
      // the flag depends on all words: written last
      dst[j] = lo;
    }
    dst[PW + j] = hi;
  }
}

__global__ void fetch_kernel(fc::GenomeView g, int64_t gp, int64_t n, char* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = fc::gcode(g, gp + i);
  out[i] = "ACGTN"[c];
}

void genome_free(fc_genome& g) {
  if (g.shared) {
    g = fc_genome();
    return;
  }
  cudaFree(g.d_plo);
  cudaFree(g.d_phi);
  cudaFree(g.d_pn);
  cudaFree(g.d_tiles);
  cudaFree(g.d_off);
  cudaFree(g.d_size);
  g = fc_genome();
}

int genome_layout(fc_ctx* ctx) {
  fc_genome& g = ctx->genome;
  int64_t off = FC_GENOME_PAD;
  g.offs.clear();
  for (size_t i = 0; i < g.sizes.size(); ++i) {
    off = (off + ALIGN - 1) / ALIGN * ALIGN;
    g.offs.push_back(off);
    off += g.sizes[i] + FC_GENOME_PAD;
  }
  g.total = (off + ALIGN - 1) / ALIGN * ALIGN + ALIGN;
  int64_t bp = (g.total + SLACK_BASES) / 8 + 256;  // bytes per plane
  FC_CUDA(ctx, cudaMalloc(&g.d_plo, bp));
  FC_CUDA(ctx, cudaMalloc(&g.d_phi, bp));
  FC_CUDA(ctx, cudaMalloc(&g.d_pn, bp));
  FC_CUDA(ctx, cudaMemset(g.d_plo, 0, bp));
  FC_CUDA(ctx, cudaMemset(g.d_phi, 0, bp));
  FC_CUDA(ctx, cudaMemset(g.d_pn, 0xFF, bp));
  size_t nc = g.sizes.size();
  FC_CUDA(ctx, cudaMalloc(&g.d_off, sizeof(int64_t) * std::max<size_t>(nc, 1)));
  FC_CUDA(ctx, cudaMalloc(&g.d_size, sizeof(int64_t) * std::max<size_t>(nc, 1)));
  FC_CUDA(ctx, cudaMemcpy(g.d_off, g.offs.data(), sizeof(int64_t) * nc, cudaMemcpyHostToDevice));
  FC_CUDA(ctx, cudaMemcpy(g.d_size, g.sizes.data(), sizeof(int64_t) * nc, cudaMemcpyHostToDevice));
  g.dev_bytes = 3 * bp + 2 * sizeof(int64_t) * nc;
  return FC_OK;
}

struct Packer {
  fc_ctx* ctx;
  uint8_t* h_stage[2] = {nullptr, nullptr};
  uint8_t* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2];
  unsigned long long* d_cnt = nullptr;
  int cur = 0;
  int64_t fill = 0;       // bytes in h_stage[cur]
  int64_t chrom = -1;     // chromosome being filled
  int64_t chrom_pos = 0;  // local base index of h_stage[cur][0]

  int init() {
    for (int k = 0; k < 2; ++k) {
      FC_CUDA(ctx, cudaMallocHost(&h_stage[k], CHUNK));
      FC_CUDA(ctx, cudaMalloc(&d_stage[k], CHUNK));
      FC_CUDA(ctx, cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    }
    FC_CUDA(ctx, cudaMalloc(&d_cnt, 2 * sizeof(unsigned long long)));
    FC_CUDA(ctx, cudaMemset(d_cnt, 0, 2 * sizeof(unsigned long long)));
    return FC_OK;
  }
  void fini() {
    for (int k = 0; k < 2; ++k) {
      if (h_stage[k]) cudaFreeHost(h_stage[k]);
      if (d_stage[k]) cudaFree(d_stage[k]);
      cudaEventDestroy(done[k]);
    }
    cudaFree(d_cnt);
  }
  // push the current staging buffer: bases [chrom_pos, chrom_pos+fill) of chromosome `chrom`
  int flush(bool last_of_chrom) {
    fc_genome& g = ctx->genome;
    if (chrom < 0) return FC_OK;
    if (fill == 0 && !last_of_chrom) return FC_OK;
    int64_t gbase = g.offs[chrom] + chrom_pos;
    // a partial trailing word is completed with N by the kernel; only the last chunk of a chromosome may be partial
    int64_t n_words = (fill + 31) / 32;
    if (n_words > 0) {
      cudaStream_t st = ctx->own_stream;
      FC_CUDA(ctx, cudaMemcpyAsync(d_stage[cur], h_stage[cur], fill, cudaMemcpyHostToDevice, st));
      int threads = 256;
      int64_t nthreads = (n_words + 31) / 32 * 32;
      int blocks = (int)((nthreads + threads - 1) / threads);
      pack_chunk_kernel<<<blocks, threads, 0, st>>>(d_stage[cur], fill, gbase, n_words, g.d_plo, g.d_phi, g.d_pn, d_cnt);
      FC_LAUNCH_CHECK(ctx);
      FC_CUDA(ctx, cudaEventRecord(done[cur], st));
    }
    chrom_pos += fill;
    fill = 0;
    cur ^= 1;
    FC_CUDA(ctx, cudaEventSynchronize(done[cur]));  // the buffer we are about to reuse
    return FC_OK;
  }
  int begin_chrom(int64_t c) {
    int rc = flush(true);
    if (rc) return rc;
    chrom = c;
    chrom_pos = 0;
    return FC_OK;
  }
  int append(const uint8_t* p, int64_t n) {
    while (n > 0) {
      int64_t k = std::min(n, CHUNK - fill);
      memcpy(h_stage[cur] + fill, p, k);
      fill += k;
      p += k;
      n -= k;
      if (fill == CHUNK) {
        int rc = flush(false);
        if (rc) return rc;
      }
    }
    return FC_OK;
  }
  int finish() {
    int rc = flush(true);
    if (rc) return rc;
    FC_CUDA(ctx, cudaStreamSynchronize(ctx->own_stream));
    unsigned long long h[2];
    FC_CUDA(ctx, cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
    ctx->genome.n_n = (int64_t)h[0];
    ctx->genome.n_other = (int64_t)h[1];
    return FC_OK;
  }
};

}  // namespace

extern "C" int fc_genome_load_fasta(fc_ctx* ctx, const char* path) {
  if (!ctx || !path) return FC_E_ARG;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  FILE* fh = fopen(path, "rb");
  if (!fh) return fc_fail(ctx, FC_E_IO, "cannot open genome FASTA '%s'", path);
  genome_free(ctx->genome);
  fc_genome& g = ctx->genome;

  // pass 1: names and sizes (find_circ.py:120-155: name = first word after '>', size = sum of stripped line lengths)
  std::vector<char> line_buf(1 << 20);
  auto strip_len = [](const char* s, size_t n, size_t& lead) {
    lead = 0;
    while (lead < n && (s[lead] == ' ' || s[lead] == '\t' || s[lead] == '\r' || s[lead] == '\n' || s[lead] == '\f' || s[lead] == '\v')) lead++;
    while (n > lead && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\r' || s[n - 1] == '\n' || s[n - 1] == '\f' || s[n - 1] == '\v')) n--;
    return n - lead;
  };
  {
    int64_t cur = -1;
    bool continued = false;  // previous fgets did not reach the newline
    bool header_cont = false;
    while (fgets(line_buf.data(), (int)line_buf.size(), fh)) {
      size_t n = strlen(line_buf.data());
      bool complete = n > 0 && line_buf[n - 1] == '\n';
      if (!continued && line_buf[0] == '>') {
        std::string nm(line_buf.data() + 1, n - 1);
        size_t a = nm.find_first_not_of(" \t\r\n");
        if (a == std::string::npos) {
          fclose(fh);
          return fc_fail(ctx, FC_E_FORMAT, "empty FASTA header in '%s'", path);
        }
        size_t b = nm.find_first_of(" \t\r\n", a);
        g.names.push_back(nm.substr(a, b == std::string::npos ? std::string::npos : b - a));
        g.sizes.push_back(0);
        cur = (int64_t)g.sizes.size() - 1;
        header_cont = !complete;
      } else if (continued && header_cont) {
        header_cont = !complete;
      } else {
        if (cur < 0) {
          size_t lead;
          if (strip_len(line_buf.data(), n, lead) == 0) {
            continued = !complete;
            continue;
          }
          fclose(fh);
          return fc_fail(ctx, FC_E_FORMAT, "sequence data before the first '>' in '%s'", path);
        }
        size_t lead;
        g.sizes[cur] += (int64_t)strip_len(line_buf.data(), n, lead);
      }
      continued = !complete;
    }
  }
  if (g.names.empty()) {
    fclose(fh);
    return fc_fail(ctx, FC_E_FORMAT, "no sequences in '%s'", path);
  }
  for (size_t i = 0; i < g.sizes.size(); ++i) g.n_bases += g.sizes[i];
  int rc = genome_layout(ctx);
  if (rc) {
    fclose(fh);
    return rc;
  }

  // pass 2: stream the letters through pinned staging -> device packer
  Packer pk;
  pk.ctx = ctx;
  rc = pk.init();
  if (rc == FC_OK) {
    rewind(fh);
    int64_t cur = -1;
    bool continued = false, header_cont = false;
    while (rc == FC_OK && fgets(line_buf.data(), (int)line_buf.size(), fh)) {
      size_t n = strlen(line_buf.data());
      bool complete = n > 0 && line_buf[n - 1] == '\n';
      if (!continued && line_buf[0] == '>') {
        cur++;
        rc = pk.begin_chrom(cur);
        header_cont = !complete;
      } else if (continued && header_cont) {
        header_cont = !complete;
      } else if (cur >= 0) {
        size_t lead;
        size_t m = strip_len(line_buf.data(), n, lead);
        if (m) rc = pk.append(reinterpret_cast<const uint8_t*>(line_buf.data()) + lead, (int64_t)m);
      }
      continued = !complete;
    }
    if (rc == FC_OK) rc = pk.finish();
  }
  pk.fini();
  fclose(fh);
  if (rc) return rc;
  g.loaded = true;
  return FC_OK;
}

extern "C" int fc_genome_load_ascii(fc_ctx* ctx, int32_t n_chrom, const char* const* names, const uint8_t* const* seqs,
                                    const int64_t* sizes) {
  if (!ctx || n_chrom <= 0 || !names || !seqs || !sizes) return FC_E_ARG;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  genome_free(ctx->genome);
  fc_genome& g = ctx->genome;
  for (int32_t i = 0; i < n_chrom; ++i) {
    g.names.push_back(names[i]);
    g.sizes.push_back(sizes[i]);
    g.n_bases += sizes[i];
  }
  int rc = genome_layout(ctx);
  if (rc) return rc;
  Packer pk;
  pk.ctx = ctx;
  rc = pk.init();
  for (int32_t i = 0; rc == FC_OK && i < n_chrom; ++i) {
    rc = pk.begin_chrom(i);
    if (rc == FC_OK) rc = pk.append(seqs[i], sizes[i]);
  }
  if (rc == FC_OK) rc = pk.finish();
  pk.fini();
  if (rc) return rc;
  g.loaded = true;
  return FC_OK;
}

extern "C" int fc_genome_n_chrom(fc_ctx* ctx) { return ctx && ctx->genome.loaded ? (int)ctx->genome.names.size() : 0; }

extern "C" int fc_genome_chrom_name(fc_ctx* ctx, int32_t i, char* buf, int32_t cap) {
  if (!ctx || !buf || cap <= 0 || i < 0 || i >= (int32_t)ctx->genome.names.size()) return FC_E_ARG;
  snprintf(buf, cap, "%s", ctx->genome.names[i].c_str());
  return FC_OK;
}

extern "C" int64_t fc_genome_chrom_size(fc_ctx* ctx, int32_t i) {
  if (!ctx || i < 0 || i >= (int32_t)ctx->genome.sizes.size()) return FC_E_ARG;
  return ctx->genome.sizes[i];
}

extern "C" int64_t fc_genome_chrom_offset(fc_ctx* ctx, int32_t i) {
  if (!ctx || !ctx->genome.loaded || i < 0 || i >= (int32_t)ctx->genome.offs.size()) return -1;
  return ctx->genome.offs[i];
}

extern "C" int fc_genome_chrom_id(fc_ctx* ctx, const char* name) {
  if (!ctx || !name) return -1;
  for (size_t i = 0; i < ctx->genome.names.size(); ++i)
    if (ctx->genome.names[i] == name) return (int)i;
  return -1;
}

extern "C" int fc_genome_stats(fc_ctx* ctx, int64_t stats[4]) {
  if (!ctx || !stats) return FC_E_ARG;
  stats[0] = ctx->genome.n_bases;
  stats[1] = ctx->genome.n_n;
  stats[2] = ctx->genome.n_other;
  stats[3] = ctx->genome.dev_bytes;
  return FC_OK;
}

extern "C" int fc_genome_fetch(fc_ctx* ctx, int32_t chrom, int64_t start, int64_t end, char* h_out) {
  if (!ctx || !h_out) return FC_E_ARG;
  fc_genome& g = ctx->genome;
  if (!g.loaded) return fc_fail(ctx, FC_E_NOGENOME, "no genome loaded");
  if (chrom < 0 || chrom >= (int32_t)g.names.size()) return fc_fail(ctx, FC_E_ARG, "unknown chromosome id %d", chrom);
  int64_t n = end - start;
  if (n <= 0) return FC_OK;
  if (start < -(int64_t)FC_GENOME_PAD || end > g.sizes[chrom] + FC_GENOME_PAD)
    return fc_fail(ctx, FC_E_RANGE, "fetch [%lld,%lld) exceeds the padding around chromosome %d", (long long)start,
                   (long long)end, chrom);
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  char* d = nullptr;
  FC_CUDA(ctx, cudaMalloc(&d, n));
  int threads = 256;
  fetch_kernel<<<(int)((n + threads - 1) / threads), threads, 0, ctx->own_stream>>>(g.view(), g.offs[chrom] + start, n, d);
  FC_LAUNCH_CHECK(ctx);
  cudaError_t e = cudaMemcpyAsync(h_out, d, n, cudaMemcpyDeviceToHost, ctx->own_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->own_stream);
  cudaFree(d);
  if (e != cudaSuccess) return fc_fail(ctx, FC_E_CUDA, "fetch copy failed: %s", cudaGetErrorString(e));
  return FC_OK;
}

// (re)build the tile store so that windows of up to `w` bases fit inside one tile
int fc_genome_ensure_tiles(fc_ctx* ctx, int w, cudaStream_t st) {
  fc_genome& g = ctx->genome;
  if (!g.loaded) return fc_fail(ctx, FC_E_NOGENOME, "no genome loaded");
  if (w < 8) w = 8;
  if (w > 256) w = 256;
  if (g.d_tiles && w <= g.tile_W) return FC_OK;
  if (g.shared) return fc_fail(ctx, FC_E_STATE, "a shared genome store cannot rebuild its tiles for %d-base windows: scan on the owning context first", w);
  int T, P, cap;
  fc::tile_geometry(w, T, P, cap);
  const int S = fc::TILE_STRIDE;
  int64_t n_tiles = g.total / S + 2;
  size_t bytes = (size_t)n_tiles * 32 * T + 256;
  if (g.d_tiles) {
    FC_CUDA(ctx, cudaStreamSynchronize(st));
    FC_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(g.d_tiles);
    g.d_tiles = nullptr;
    g.dev_bytes -= g.tile_bytes;
  }
  FC_CUDA(ctx, cudaMalloc(&g.d_tiles, bytes));
  int threads = 128;
  build_tiles_kernel<<<(unsigned)((n_tiles + threads - 1) / threads), threads, 0, st>>>(g.d_plo, g.d_phi, g.d_pn, n_tiles, T,
                                                                                       S, g.d_tiles);
  FC_LAUNCH_CHECK(ctx);
  FC_CUDA(ctx, cudaStreamSynchronize(st));  // (rare) scans on other streams may follow at once
  g.tile_T = T;
  g.tile_S = S;
  g.tile_W = cap;
  g.tile_bytes = (int64_t)bytes;
  g.dev_bytes += g.tile_bytes;
  return FC_OK;
}

void fc_genome_release(fc_ctx* ctx) { genome_free(ctx->genome); }

// a second context of the same process and device looks at the store of `src` (no copy); src must outlive dst
extern "C" int fc_genome_share(fc_ctx* dst, fc_ctx* src) {
  if (!dst || !src || dst == src) return FC_E_ARG;
  if (!src->genome.loaded) return fc_fail(dst, FC_E_NOGENOME, "the source context holds no genome");
  if (dst->device != src->device) return fc_fail(dst, FC_E_ARG, "contexts live on different devices");
  genome_free(dst->genome);
  dst->genome = src->genome;
  dst->genome.shared = true;
  return FC_OK;
}
