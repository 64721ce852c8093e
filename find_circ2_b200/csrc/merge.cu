// merge.cu -- keyed merge of several junction tables on the device: the arithmetic of merge_bed.py
// (/root/reference/merge_bed.py:88-142, column map :117-130).  Rows of all input tables that share
// (chrom, start, end, strand) form one output row; numeric columns are summed, maximised or minimised over the rows of a
// group in input order (the order merge_bed.py walks its `support` lists: file by file), and a bit mask tells which inputs
// support the group.  Text columns (names, sample lists, keyword sets) stay with the host (find_circ2_b200/merge_bed.py).
//
// Same machinery as the sort-based junction aggregation (agg.cu): stable CUB radix sorts put the rows in key order, head
// flags + a scan number the groups, one thread per group reduces its rows sequentially -- deterministic, bit-identical to a
// python loop over the same rows.
#include <cub/cub.cuh>

#include "fc_internal.cuh"

namespace {

__global__ void merge_keys_kernel(int64_t n, const uint32_t* __restrict__ chrom, const int32_t* __restrict__ start,
                                  const int32_t* __restrict__ end, const uint8_t* __restrict__ strand, uint64_t* __restrict__ k1,
                                  uint64_t* __restrict__ k2, uint32_t* __restrict__ iota) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  k1[i] = ((uint64_t)chrom[i] << 32) | (uint32_t)((uint32_t)start[i] ^ 0x80000000u);
  k2[i] = ((uint64_t)((uint32_t)end[i] ^ 0x80000000u) << 8) | strand[i];
  iota[i] = (uint32_t)i;
}

__global__ void merge_gather_key_kernel(int64_t n, const uint64_t* __restrict__ k, const uint32_t* __restrict__ perm, uint64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = k[perm[i]];
}

__global__ void merge_heads_kernel(int64_t n, const uint64_t* __restrict__ k1, const uint64_t* __restrict__ k2,
                                   const uint32_t* __restrict__ perm, uint32_t* __restrict__ head) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || k1[perm[i]] != k1[perm[i - 1]] || k2[perm[i]] != k2[perm[i - 1]]) ? 1u : 0u;
}

__global__ void merge_starts_kernel(int64_t n, const uint32_t* __restrict__ head, const uint32_t* __restrict__ gid_incl,
                                    const uint32_t* __restrict__ perm, uint32_t* __restrict__ start, uint32_t* __restrict__ group_of_row) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t g = gid_incl[i] - 1u;
  if (head[i]) start[g] = (uint32_t)i;
  group_of_row[perm[i]] = g;
}

// one thread per group: its rows in input order
__global__ void merge_reduce_kernel(int64_t n, int64_t n_groups, const uint32_t* __restrict__ start, const uint32_t* __restrict__ perm,
                                    const uint8_t* __restrict__ src, int n_cols, const double* __restrict__ vals,
                                    const uint8_t* __restrict__ op, unsigned long long* __restrict__ support, double* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const int64_t lo = start[g], hi = g + 1 < n_groups ? (int64_t)start[g + 1] : n;
  unsigned long long mask = 0ull;
  for (int64_t i = lo; i < hi; ++i) mask |= 1ull << (src[perm[i]] & 63);
  support[g] = mask;
  for (int c = 0; c < n_cols; ++c) {
    const double* col = vals + (int64_t)c * n;
    double acc = col[perm[lo]];
    for (int64_t i = lo + 1; i < hi; ++i) {
      const double v = col[perm[i]];
      acc = op[c] == 0 ? acc + v : (op[c] == 1 ? (v > acc ? v : acc) : (v < acc ? v : acc));
    }
    out[(int64_t)c * n_groups + g] = acc;
  }
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" int fc_merge_tables(fc_ctx* ctx, int64_t n, const uint32_t* h_chrom, const int32_t* h_start, const int32_t* h_end,
                               const uint8_t* h_strand, const uint8_t* h_src, int32_t n_cols, const double* h_vals, const uint8_t* h_op,
                               int64_t* out_n_groups, uint32_t* h_group_of_row, uint64_t* h_support, double* h_out) {
  if (!ctx || n < 0 || n >= (1ll << 31) || n_cols < 0 || n_cols > 64 || !out_n_groups) return FC_E_ARG;
  *out_n_groups = 0;
  if (n == 0) return FC_OK;
  if (!h_chrom || !h_start || !h_end || !h_strand || !h_src || !h_group_of_row || !h_support || (n_cols && (!h_vals || !h_op || !h_out))) return FC_E_ARG;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  const size_t N = (size_t)n;
  fc_dbuf b[14];
  struct Free {
    fc_dbuf* b;
    ~Free() {
      for (int k = 0; k < 14; ++k) b[k].release();
    }
  } guard{b};
  const size_t sizes[14] = {4 * N, 4 * N, 4 * N, N, N, 8 * N * (size_t)(n_cols ? n_cols : 1), 8 * N, 8 * N, 8 * N, 4 * N, 4 * N, 4 * N, 4 * N, 64};
  for (int k = 0; k < 14; ++k) FC_CUDA(ctx, b[k].reserve(sizes[k], st, false, 0));
  uint32_t* d_chrom = (uint32_t*)b[0].p;
  int32_t *d_start = (int32_t*)b[1].p, *d_end = (int32_t*)b[2].p;
  uint8_t *d_strand = (uint8_t*)b[3].p, *d_src = (uint8_t*)b[4].p;
  double* d_vals = (double*)b[5].p;
  uint64_t *k1 = (uint64_t*)b[6].p, *k2 = (uint64_t*)b[7].p, *ks = (uint64_t*)b[8].p;
  uint32_t *pa = (uint32_t*)b[9].p, *pb = (uint32_t*)b[10].p, *head = (uint32_t*)b[11].p, *gid = (uint32_t*)b[12].p;
  uint8_t* d_op = (uint8_t*)b[13].p;
  FC_CUDA(ctx, cudaMemcpyAsync(d_chrom, h_chrom, 4 * N, cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_start, h_start, 4 * N, cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_end, h_end, 4 * N, cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_strand, h_strand, N, cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_src, h_src, N, cudaMemcpyHostToDevice, st));
  if (n_cols) {
    FC_CUDA(ctx, cudaMemcpyAsync(d_vals, h_vals, 8 * N * (size_t)n_cols, cudaMemcpyHostToDevice, st));
    FC_CUDA(ctx, cudaMemcpyAsync(d_op, h_op, (size_t)n_cols, cudaMemcpyHostToDevice, st));
  }
  merge_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(n, d_chrom, d_start, d_end, d_strand, k1, k2, pa);
  FC_LAUNCH_CHECK(ctx);
  // lexicographic order by (k1, k2): stable sort by the minor key, then by the major key
  fc_dbuf tmpbuf, kb;
  FC_CUDA(ctx, kb.reserve(8 * N, st, false, 0));
  auto sort_by = [&](const uint64_t* key, uint32_t* vin, uint32_t* vout) -> int {
    merge_gather_key_kernel<<<nblk(n, 256), 256, 0, st>>>(n, key, vin, ks);
    FC_LAUNCH_CHECK(ctx);
    size_t tmp = 0;
    FC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, ks, (uint64_t*)kb.p, vin, vout, n, 0, 64, st));
    FC_CUDA(ctx, tmpbuf.reserve(tmp, st, false, 0));
    FC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmpbuf.p, tmp, ks, (uint64_t*)kb.p, vin, vout, n, 0, 64, st));
    ctx->launches += 9;
    return FC_OK;
  };
  int rc = sort_by(k2, pa, pb);
  if (!rc) rc = sort_by(k1, pb, pa);
  if (rc) {
    tmpbuf.release();
    kb.release();
    return rc;
  }
  uint32_t* perm = pa;
  merge_heads_kernel<<<nblk(n, 256), 256, 0, st>>>(n, k1, k2, perm, head);
  FC_LAUNCH_CHECK(ctx);
  {
    size_t tmp = 0;
    FC_CUDA(ctx, cub::DeviceScan::InclusiveSum(nullptr, tmp, head, gid, n, st));
    FC_CUDA(ctx, tmpbuf.reserve(tmp, st, false, 0));
    FC_CUDA(ctx, cub::DeviceScan::InclusiveSum(tmpbuf.p, tmp, head, gid, n, st));
    ctx->launches += 2;
  }
  uint32_t ng32 = 0;
  FC_CUDA(ctx, cudaMemcpyAsync(&ng32, gid + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  const int64_t ng = ng32;
  fc_dbuf gstart, grow, gsup, gout;
  cudaError_t e = gstart.reserve(4 * (size_t)ng, st, false, 0);
  if (e == cudaSuccess) e = grow.reserve(4 * N, st, false, 0);
  if (e == cudaSuccess) e = gsup.reserve(8 * (size_t)ng, st, false, 0);
  if (e == cudaSuccess) e = gout.reserve(8 * (size_t)ng * (size_t)(n_cols ? n_cols : 1), st, false, 0);
  if (e == cudaSuccess) {
    merge_starts_kernel<<<nblk(n, 256), 256, 0, st>>>(n, head, gid, perm, (uint32_t*)gstart.p, (uint32_t*)grow.p);
    merge_reduce_kernel<<<nblk(ng, 128), 128, 0, st>>>(n, ng, (const uint32_t*)gstart.p, perm, d_src, n_cols, d_vals, d_op,
                                                      (unsigned long long*)gsup.p, (double*)gout.p);
    ctx->launches += 2;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_group_of_row, grow.p, 4 * N, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_support, gsup.p, 8 * (size_t)ng, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && n_cols) e = cudaMemcpyAsync(h_out, gout.p, 8 * (size_t)ng * (size_t)n_cols, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  gstart.release();
  grow.release();
  gsup.release();
  gout.release();
  tmpbuf.release();
  kb.release();
  if (e != cudaSuccess) return fc_fail(ctx, FC_E_CUDA, "fc_merge_tables: %s", cudaGetErrorString(e));
  *out_n_groups = ng;
  return FC_OK;
}
