// scan.cu -- kernels around scan_core.cuh: read packing, the breakpoint scan, the tie enumerator, and the
// host-buffer convenience entry (H2D + pack + scan + D2H).
//
// Replaces JunctionSpan.find_breakpoints (/root/reference/find_circ.py:854-974); see scan_core.cuh.
#include <stdlib.h>

#include <vector>

#include "fc_internal.cuh"

namespace {

// ---------------------------------------------------------------- ASCII -> bit planes of the internal read part
// one thread per pair; rows are read 16 bytes at a time when the matrix is 16-byte aligned
__global__ void pack_reads_kernel(int64_t n, const uint8_t* __restrict__ ascii, int32_t stride,
                                  const int32_t* __restrict__ l, int32_t n_words, uint32_t* __restrict__ rlo,
                                  uint32_t* __restrict__ rhi, uint32_t* __restrict__ rn, uint8_t* __restrict__ flags,
                                  int aligned16) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int li = l[i];
  const uint8_t* row = ascii + i * (int64_t)stride;
  uint32_t any = 0;
  for (int w = 0; w < n_words; ++w) {
    int count = li - 32 * w;
    count = count < 0 ? 0 : (count > 32 ? 32 : count);
    uint32_t lo = 0, hi = 0, nn = 0;
    if (count > 0) {
      if (aligned16) {
        uint4 v[2];
        v[0] = __ldg(reinterpret_cast<const uint4*>(row + 32 * w));
        v[1] = count > 16 ? __ldg(reinterpret_cast<const uint4*>(row + 32 * w + 16)) : make_uint4(0, 0, 0, 0);
        fc::pack32(reinterpret_cast<const uint8_t*>(v), count, lo, hi, nn);
      } else {
        fc::pack32(row + 32 * w, count, lo, hi, nn);
      }
    }
    rlo[(int64_t)w * n + i] = lo;
    rhi[(int64_t)w * n + i] = hi;
    rn[(int64_t)w * n + i] = nn;
    any |= nn;
  }
  if (any) flags[i] |= (uint8_t)FC_PF_READ_N;
}

// ---------------------------------------------------------------- SoA batch -> packed batch (fc_batch)
// One thread per pair.  The 16-byte descriptor carries the GLOBAL coordinates of both windows (so that the scan needs no
// chromosome table before it can fetch its tiles) and the validity verdict of find_circ.py:194-211 (both windows must
// touch the chromosome); the read planes go from word-major columns to one row per pair (one vector load in the scan).
__global__ void pack_batch_kernel(fc::GenomeView g, int64_t n, const int32_t* __restrict__ chrom, const int32_t* __restrict__ a_start,
                                  const int32_t* __restrict__ b_end, const int32_t* __restrict__ l, const uint8_t* __restrict__ flags,
                                  const uint8_t* __restrict__ wden, const uint8_t* __restrict__ frag, fc::ReadView rv, int nw_out,
                                  uint4* __restrict__ meta, uint32_t* __restrict__ reads, uint32_t* __restrict__ rn_out,
                                  const int16_t* __restrict__ q_a, const int16_t* __restrict__ q_b, uint32_t* __restrict__ q_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t c = chrom[i], li = l[i];
  uint32_t fl = flags[i] & 7u;
  int64_t ga = 0, gb = 0;
  bool ok = false;
  if (li >= 0 && c >= 0 && c < g.n_chrom && li + 2 <= g.pad) {
    const int64_t size = g.chrom_size[c];
    const int w = li + 2;
    const int64_t as = a_start[i], be = b_end[i];
    ok = (uint64_t)(as + w) <= (uint64_t)(size + w) && (uint64_t)be <= (uint64_t)(size + w);
    if (ok) {
      const int64_t off = g.chrom_off[c];
      ga = off + as;
      gb = off + be - w;
    }
  }
  if (!ok) fl |= FC_META_INVALID | (li < 0 ? FC_META_NEGATIVE : 0u);
  const uint32_t fr = frag ? (uint32_t)(frag[i] & 15u) : 15u;  // 15 = (3,3): nothing known about the fragment's other rows
  const uint32_t lfield = ok ? (uint32_t)li : 0u;
  meta[i] = make_uint4((uint32_t)ga, (uint32_t)gb,
                       (uint32_t)(ga >> 32) | ((uint32_t)(gb >> 32) << 6) | (lfield << 12) | (fl << 24) | (fr << 28),
                       ((uint32_t)c & 0xFFFFFFu) | ((uint32_t)(wden ? wden[i] : 1u) << 24));
  uint32_t* row = reads + i * (int64_t)(2 * nw_out);
  for (int k = 0; k < nw_out; ++k) {
    const bool have = k < rv.n_words;
    row[k] = have ? rv.rlo[(int64_t)k * rv.stride + i * rv.pair_stride] : 0u;
    row[nw_out + k] = have ? rv.rhi[(int64_t)k * rv.stride + i * rv.pair_stride] : 0u;
  }
  if ((fl & FC_PF_READ_N) && rn_out)
    for (int k = 0; k < nw_out; ++k)
      rn_out[i * (int64_t)nw_out + k] = k < rv.n_words ? rv.rn[(int64_t)k * rv.stride + i * rv.rn_pair_stride] : 0u;
  if (q_out) q_out[i] = (uint32_t)(uint16_t)q_a[i] | ((uint32_t)(uint16_t)q_b[i] << 16);
}

// ---------------------------------------------------------------- the scan
struct BatchView {
  const uint4* meta;
  const uint32_t* reads;  // rows of 2 * nw words
  const uint32_t* rn;     // rows of nw words (pairs flagged READ_N only)
  int64_t n;
  int32_t nw;
};
struct Payload {  // what becomes of an accepted pair (emit_core.cuh)
  const uint32_t* q;  // q_a | q_b << 16
  const uint64_t* read_hash;
  const uint64_t* qname_hash;
  uint64_t idx_base;
  const uint64_t* idx;
  unsigned long long* n_recs;
  fc_jrec* recs;
  unsigned long long* no_name;  // counts records that have neither a name hash nor fragment fields (fc_agg_finalize then fails)
};

// the read planes of pair i: one vector load per pair when the row is 8 or 16 bytes, two or four for 32 / 64 bytes
// (nw is uniform over the launch: the branches do not diverge)
template <int NP>
__device__ __forceinline__ void load_read_row(const uint32_t* __restrict__ reads, int64_t i, int nw, uint32_t (&rlo)[NP],
                                              uint32_t (&rhi)[NP]) {
#pragma unroll
  for (int k = 0; k < NP; ++k) rlo[k] = rhi[k] = 0u;
  if (nw == 2) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(reads) + i);
    rlo[0] = v.x;
    rhi[0] = v.z;
    if constexpr (NP >= 2) {
      rlo[1] = v.y;
      rhi[1] = v.w;
    }
  } else if (nw == 1) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(reads) + i);
    rlo[0] = v.x;
    rhi[0] = v.y;
  } else if (nw == 4) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(reads) + 2 * i), c = __ldg(reinterpret_cast<const uint4*>(reads) + 2 * i + 1);
    const uint32_t lo[4] = {a.x, a.y, a.z, a.w}, hi[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < (NP < 4 ? NP : 4); ++k) {
      rlo[k] = lo[k];
      rhi[k] = hi[k];
    }
  } else {  // any other row length: scalar loads
    const uint32_t* row = reads + i * (int64_t)(2 * nw);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      if (k < nw) {
        rlo[k] = __ldg(row + k);
        rhi[k] = __ldg(row + nw + k);
      }
    }
  }
}

// MODE 0: hits only; 1: every pair that found a breakpoint becomes a junction record of this context on the way (the hits
// do not travel to HBM and back before they are turned into records, and one launch less stands in the step); 2: the
// records go to the ranks that own their junction keys (multi-GPU)
template <int NP, int T, int BS, int MB, int MODE>
__global__ void __launch_bounds__(BS, MB) scan_kernel(fc::GenomeView g, fc::ScanCfg cfg, BatchView b, fc_hit* __restrict__ out,
                                                   Payload e, fc::P2PView pv, unsigned long long* overflow) {
  const int64_t i = (int64_t)blockIdx.x * BS + threadIdx.x;  // one pair per thread: the whole CTA reaches the emit
  fc::HitOut h;
  h.start = h.end = 0;
  h.w2 = h.w3 = 0u;
  uint4 m = make_uint4(0u, 0u, 0u, 0u);
  if (i < b.n) {
    m = __ldg(b.meta + i);
    uint32_t rlo[NP], rhi[NP];
    load_read_row<NP>(b.reads, i, b.nw, rlo, rhi);
    if constexpr (MODE != 0) {
      // the payload of the record is only looked at after the scan: bring it into L2 now, beside the tile loads
      asm volatile("prefetch.global.L2 [%0];" ::"l"(e.read_hash + i));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(e.q + i));
      if (e.qname_hash) asm volatile("prefetch.global.L2 [%0];" ::"l"(e.qname_hash + i));
    }
    const int64_t ga = (int64_t)m.x | ((int64_t)(m.z & 63u) << 32), gb = (int64_t)m.y | ((int64_t)((m.z >> 6) & 63u) << 32);
    const int l = (int)((m.z >> 12) & 0xFFFu);
    const uint32_t fl = (m.z >> 24) & 15u;
    const bool backsplice = fl & 1u, minus_span = fl & 2u, read_n = fl & 4u;
    fc::Best best;
    best.init();
    uint32_t extra = 0;
    int64_t off = 0;
    if (!(fl & FC_META_INVALID)) {
      off = g.chrom_off[m.w & 0xFFFFFFu];  // (needed only to print the hit: in flight beside the tile loads)
      fc::ReadView rv{b.reads, b.reads + b.nw, b.rn, b.n, b.nw, 1, 2 * (int64_t)b.nw, (int64_t)b.nw};
      if (cfg.noncanonical || (l + 2 > 32 * NP)) {
        extra |= fc::W3_SLOW;
        fc::NoEmit ne;
        fc::scan_per_base(g, cfg, ga, gb, l, minus_span, rv, i, read_n, best, ne);
      } else {
        fc::scan_fast<NP, T>(g, cfg, ga, gb, l, minus_span, read_n, rlo, rhi, rv, i, best);
      }
    } else if (!(fl & FC_META_NEGATIVE)) {
      extra |= fc::W3_RANGE;
    }
    fc::finish(best, (int32_t)(ga - off), (int32_t)(gb + l + 2 - off), l, backsplice, extra, h);
    reinterpret_cast<uint4*>(out)[i] = make_uint4((uint32_t)h.start, (uint32_t)h.end, h.w2, h.w3);
  }
  if constexpr (MODE != 0) {
    const bool accept = (h.w2 & 0xFFFFu) != 0u;
    // Is this the first record of its FRAGMENT for its junction (n_frags = distinct read names per junction,
    // find_circ.py:584-586)?  The rows of a fragment are neighbours and the descriptor says how many come before and after:
    // when they all sit in this warp the answer is a few shuffles, and the aggregation needs no name set for the record.
    const uint32_t fr = m.z >> 28, back = fr & 3u, fwd = fr >> 2;
    const unsigned lane = threadIdx.x & 31;
    const bool known = fr != 15u && lane >= back && lane + fwd < 32u;
    bool dup = false;
    if (__any_sync(0xffffffffu, known && back != 0u)) {
      const uint32_t kw = (h.w3 & 1u) | (((m.z >> 24) & 1u) << 1) | ((m.w & 0xFFFFFFu) << 2) | ((uint32_t)accept << 31);
#pragma unroll
      for (int k = 1; k <= 3; ++k) {
        const uint32_t okw = __shfl_up_sync(0xffffffffu, kw, k);
        const int32_t os = __shfl_up_sync(0xffffffffu, h.start, k), oe = __shfl_up_sync(0xffffffffu, h.end, k);
        if ((uint32_t)k <= back && known && accept && okw == kw && os == h.start && oe == h.end) dup = true;
      }
    }
    fc_jrec r;
    if (accept) {
      const uint32_t q = e.q[i];
      // a batch without name hashes: the rows of a fragment share the stream position of its first row as their name
      // (read names are unique per fragment on this path, so that is as good as a hash of the name)
      const int64_t first_row = i - (int64_t)(fr != 15u ? back : 0u);
      const uint64_t qh = e.qname_hash ? e.qname_hash[i]
                                       : fc_mix64((e.idx ? e.idx[first_row < 0 ? 0 : first_row] : e.idx_base + (uint64_t)first_row) ^ 0x6a09e667f3bcc909ULL);
      r = fc::make_record_from(h.start, h.end, h.w2, h.w3, m.w & 0xFFFFFFu, (m.z >> 24) & 7u, m.w >> 24, (int16_t)(q & 0xFFFFu),
                               (int16_t)(q >> 16), e.read_hash[i], qh, e.idx ? e.idx[i] : e.idx_base + (uint64_t)i);
      if (known)
        r.sk |= FC_SK_NAME_KNOWN | (dup ? FC_SK_NAME_DUP : 0u);
      else if (!e.qname_hash && fr == 15u)
        atomicAdd(e.no_name, 1ull);
    }
    if constexpr (MODE == 1)
      fc::emit_block<BS>(accept, r, e.n_recs, e.recs);
    else
      fc::emit_p2p_block<BS>(accept, r, pv, overflow);
  }
}

// ---------------------------------------------------------------- --all-hits: every tie of every pair
struct TieEmit {
  int best;
  fc_hit* dst;
  int a_start, b_end, l;
  bool backsplice;
  int n_ties;
  int k;
  __device__ void operator()(int s, int x, uint32_t strand, uint32_t sig, int dist, int ov) {
    if (s != best) return;
    fc::Best b;
    b.score = s;
    b.n_ties = n_ties;
    b.x = x;
    b.info = strand | (sig << 1) | ((uint32_t)dist << 16) | ((uint32_t)ov << 24);
    fc::HitOut h;
    fc::finish(b, a_start, b_end, l, backsplice, fc::W3_SLOW, h);
    dst[k++] = fc_hit{h.start, h.end, h.w2, h.w3};
  }
};

__global__ void ties_kernel(fc::GenomeView g, fc::ScanCfg cfg, fc::ReadView rv, const int32_t* __restrict__ chrom,
                            const int32_t* __restrict__ a_start, const int32_t* __restrict__ b_end,
                            const int32_t* __restrict__ l, const uint8_t* __restrict__ flags,
                            const fc_hit* __restrict__ hits, const int64_t* __restrict__ tie_off,
                            fc_hit* __restrict__ ties) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rv.n) return;
  int nh = (int)(hits[i].w2 & 0xFFFFu);
  if (nh == 0) return;
  if (hits[i].w3 & fc::W3_RANGE) return;
  int li = l[i];
  uint32_t fl = flags[i];
  int64_t off = g.chrom_off[chrom[i]];
  int64_t ga = off + a_start[i], gb = off + (int64_t)b_end[i] - (li + 2);
  fc::Best best;
  best.init();
  fc::NoEmit ne;
  fc::scan_per_base(g, cfg, ga, gb, li, (fl & 2u) != 0, rv, i, (fl & 4u) != 0, best, ne);
  TieEmit te;
  te.best = best.score;
  te.dst = ties + tie_off[i];
  te.a_start = a_start[i];
  te.b_end = b_end[i];
  te.l = li;
  te.backsplice = fl & 1u;
  te.n_ties = best.n_ties;
  te.k = 0;
  fc::Best dummy;
  dummy.init();
  fc::scan_per_base(g, cfg, ga, gb, li, (fl & 2u) != 0, rv, i, (fl & 4u) != 0, dummy, te);
}

int check_pairs(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pr) {
  if (!ctx || !p || !pr) return FC_E_ARG;
  if (!ctx->genome.loaded) return fc_fail(ctx, FC_E_NOGENOME, "fc_scan called before a genome was loaded");
  if (p->asize <= 0 || p->margin < 0 || p->margin >= p->asize || p->maxdist < 0 || p->maxdist > 255 || p->margin > 255)
    return fc_fail(ctx, FC_E_ARG, "bad scan parameters (asize=%d margin=%d maxdist=%d)", p->asize, p->margin, p->maxdist);
  if (pr->n < 0) return fc_fail(ctx, FC_E_ARG, "negative pair count");
  if (pr->max_l + 2 > FC_GENOME_PAD)
    return fc_fail(ctx, FC_E_ARG, "internal read length %d exceeds the genome padding (%d)", pr->max_l, FC_GENOME_PAD);
  if (pr->n > 0 && pr->n_words * 32 < pr->max_l)
    return fc_fail(ctx, FC_E_ARG, "n_words=%d too small for max_l=%d", pr->n_words, pr->max_l);
  return FC_OK;
}

}  // namespace

extern "C" int fc_pack_reads(fc_ctx* ctx, int64_t n, const uint8_t* d_ascii, int32_t stride, const int32_t* d_l,
                             int32_t n_words, uint32_t* d_rlo, uint32_t* d_rhi, uint32_t* d_rn, uint8_t* d_flags,
                             void* stream) {
  if (!ctx || n < 0 || n_words < 0) return FC_E_ARG;
  if (n == 0 || n_words == 0) return FC_OK;
  int threads = 128;
  int aligned16 = ((reinterpret_cast<uintptr_t>(d_ascii) | (uintptr_t)stride) & 15u) == 0 && n_words * 32 <= stride;
  pack_reads_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      n, d_ascii, stride, d_l, n_words, d_rlo, d_rhi, d_rn, d_flags, aligned16);
  FC_LAUNCH_CHECK(ctx);
  return FC_OK;
}

// storage words per plane and pair of a packed batch whose longest internal read part has max_l bases
static int batch_words(int max_l) {
  const int w = max_l > 0 ? (max_l + 31) / 32 : 1;
  if (w > 8) return w;
  int nw = 1;
  while (nw < w) nw <<= 1;
  return nw;
}

static int scan_launch(fc_ctx* ctx, const fc_scan_params* p, const BatchView& b, int max_l, fc_hit* d_out, int mode, const Payload& e,
                       cudaStream_t st, const fc::P2PView* pvp = nullptr, unsigned long long* overflow = nullptr) {
  const int need = max_l + 2;
  if (!p->noncanonical) {
    int rc = fc_genome_ensure_tiles(ctx, need, st);  // no-op when the tile store already covers this window size
    if (rc) return rc;
  }
  fc::ScanCfg cfg{p->margin, p->maxdist, p->noncanonical, p->strandpref};
  fc::GenomeView g = ctx->genome.view();
  fc::P2PView pv = {};
  if (pvp) pv = *pvp;
#define FC_SCAN_LAUNCH_BS(NP, T, BS, MB)                                                               \
  {                                                                                                    \
    const unsigned grid = (unsigned)((b.n + BS - 1) / BS);                                             \
    if (mode == 2)                                                                                     \
      scan_kernel<NP, T, BS, MB, 2><<<grid, BS, 0, st>>>(g, cfg, b, d_out, e, pv, overflow);           \
    else if (mode == 1)                                                                                \
      scan_kernel<NP, T, BS, MB, 1><<<grid, BS, 0, st>>>(g, cfg, b, d_out, e, pv, overflow);           \
    else                                                                                               \
      scan_kernel<NP, T, BS, MB, 0><<<grid, BS, 0, st>>>(g, cfg, b, d_out, e, pv, overflow);           \
  }
  // 256-thread CTAs, 48 registers -> 5 CTAs (40 warps) per SM.  Measured alternatives on B200 (round 1): 128- and
  // 192-thread CTAs, register caps 32/40/58, and a persistent software-pipelined variant with L2 prefetch of the next
  // pair's tiles were all equal or slower.
  // (multi-GPU: 512-thread CTAs, i.e. twice as long NVLink runs per destination, change nothing either)
#define FC_SCAN_LAUNCH(NP, T) FC_SCAN_LAUNCH_BS(NP, T, 256, 5)
  // the kernel specialisation follows the tile geometry of the store (scan_core.cuh: tile_geometry)
  switch (g.tile_T) {
    case 1:
      if (need <= 64) FC_SCAN_LAUNCH(2, 1)
      else FC_SCAN_LAUNCH(3, 1)
      break;
    case 2:
      if (need <= 128) FC_SCAN_LAUNCH(4, 2)
      else FC_SCAN_LAUNCH(8, 2)
      break;
    case 4: FC_SCAN_LAUNCH(8, 4) break;
    default: FC_SCAN_LAUNCH(8, 0) break;  // no tile store (--non-canonical): master planes / per-base path
  }
#undef FC_SCAN_LAUNCH_BS
#undef FC_SCAN_LAUNCH
  FC_LAUNCH_CHECK(ctx);
  return FC_OK;
}

// fc_pairs (struct of arrays, word-major planes) -> packed batch in the context's scratch, rows [row0, row0 + n) of a
// scratch sized for `rows` rows (the chunks of one host-buffer call share it)
static int pack_soa(fc_ctx* ctx, const fc_pairs* pr, const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b,
                    const uint8_t* d_frag, cudaStream_t st, int64_t row0, int64_t rows, BatchView* out, const uint32_t** q_out) {
  const int nw = batch_words(pr->max_l > 0 ? pr->max_l : 0);
  const bool want_q = d_q_a && d_q_b;
  const size_t R = (size_t)rows;
  FC_CUDA(ctx, ctx->pk[0].reserve(16 * R, st, false, 0));
  FC_CUDA(ctx, ctx->pk[1].reserve(8 * R * nw, st, false, 0));
  FC_CUDA(ctx, ctx->pk[2].reserve(4 * R * nw, st, false, 0));
  if (want_q) FC_CUDA(ctx, ctx->pk[3].reserve(4 * R, st, false, 0));
  uint4* meta = (uint4*)ctx->pk[0].p + row0;
  uint32_t* reads = (uint32_t*)ctx->pk[1].p + (size_t)row0 * 2 * nw;
  uint32_t* rn = (uint32_t*)ctx->pk[2].p + (size_t)row0 * nw;
  uint32_t* q = want_q ? (uint32_t*)ctx->pk[3].p + row0 : nullptr;
  fc::ReadView rv{pr->d_rlo, pr->d_rhi, pr->d_rn, pr->n, pr->n_words, pr->plane_stride > 0 ? pr->plane_stride : pr->n};
  pack_batch_kernel<<<(unsigned)((pr->n + 255) / 256), 256, 0, st>>>(ctx->genome.view(), pr->n, pr->d_chrom, pr->d_a_start, pr->d_b_end,
                                                                  pr->d_l, pr->d_flags, d_wden, d_frag, rv, nw, meta, reads, rn, d_q_a,
                                                                  d_q_b, q);
  FC_LAUNCH_CHECK(ctx);
  *out = BatchView{meta, reads, rn, pr->n, nw};
  if (q_out) *q_out = q;
  return FC_OK;
}

extern "C" int fc_scan(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pr, fc_hit* d_out, void* stream) {
  int rc = check_pairs(ctx, p, pr);
  if (rc) return rc;
  if (pr->n == 0) return FC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BatchView b;
  if ((rc = pack_soa(ctx, pr, nullptr, nullptr, nullptr, nullptr, st, 0, pr->n, &b, nullptr))) return rc;
  return scan_launch(ctx, p, b, pr->max_l, d_out, 0, Payload{}, st);
}

// (row0 / rows: see pack_soa)
static int scan_emit_soa(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pr, fc_hit* d_out, const uint8_t* d_wden,
                         const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash, const uint64_t* d_qname_hash,
                         uint64_t idx_base, const uint64_t* d_idx, cudaStream_t st, int64_t row0, int64_t rows) {
  BatchView b;
  const uint32_t* q = nullptr;
  int rc = pack_soa(ctx, pr, d_wden, d_q_a, d_q_b, nullptr, st, row0, rows, &b, &q);
  if (rc) return rc;
  fc::EmitArgs ea{d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, d_idx, nullptr, nullptr};
  if ((rc = fc_agg_emit_begin(ctx, pr->n, st, &ea))) return rc;
  Payload e{q, d_read_hash, d_qname_hash, idx_base, d_idx, ea.n_recs, ea.recs, nullptr};
  if ((rc = scan_launch(ctx, p, b, pr->max_l, d_out, 1, e, st))) return rc;
  fc_agg_emit_end(ctx, pr->n, idx_base, d_idx != nullptr);
  return FC_OK;
}

extern "C" int fc_scan_emit(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pr, fc_hit* d_out, const uint8_t* d_wden,
                            const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash,
                            const uint64_t* d_qname_hash, uint64_t idx_base, const uint64_t* d_idx, void* stream) {
  int rc = check_pairs(ctx, p, pr);
  if (rc) return rc;
  if (!d_out || !d_wden || !d_q_a || !d_q_b || !d_read_hash || !d_qname_hash) return FC_E_ARG;
  if (pr->n == 0) return FC_OK;
  return scan_emit_soa(ctx, p, pr, d_out, d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, d_idx, (cudaStream_t)stream, 0,
                       pr->n);
}

extern "C" int fc_scan_emit_p2p(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pr, fc_hit* d_out, const uint8_t* d_wden,
                                const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash,
                                const uint64_t* d_qname_hash, uint64_t idx_base, void* stream) {
  int rc = check_pairs(ctx, p, pr);
  if (rc) return rc;
  if (!d_out || !d_wden || !d_q_a || !d_q_b || !d_read_hash || !d_qname_hash) return FC_E_ARG;
  fc::P2PView pv;
  unsigned long long* overflow = nullptr;
  if ((rc = fc_agg_p2p_begin(ctx, &pv, &overflow))) return rc;
  if (pr->n == 0) {
    fc_agg_p2p_end(ctx, idx_base, 0);  // an empty shard still OWNS keys: peers may have written into this rank's buffer
    return FC_OK;
  }
  cudaStream_t st = (cudaStream_t)stream;
  BatchView b;
  const uint32_t* q = nullptr;
  if ((rc = pack_soa(ctx, pr, d_wden, d_q_a, d_q_b, nullptr, st, 0, pr->n, &b, &q))) return rc;
  Payload e{q, d_read_hash, d_qname_hash, idx_base, nullptr, nullptr, nullptr, nullptr};
  if ((rc = scan_launch(ctx, p, b, pr->max_l, d_out, 2, e, st, &pv, overflow))) return rc;
  fc_agg_p2p_end(ctx, idx_base, pr->n);
  return FC_OK;
}

// ---------------------------------------------------------------- packed batches (fc_batch): what a native ingest hands over
static int check_batch(fc_ctx* ctx, const fc_scan_params* p, const fc_batch* b) {
  if (!ctx || !p || !b) return FC_E_ARG;
  if (!ctx->genome.loaded) return fc_fail(ctx, FC_E_NOGENOME, "scan called before a genome was loaded");
  if (p->asize <= 0 || p->margin < 0 || p->margin >= p->asize || p->maxdist < 0 || p->maxdist > 255 || p->margin > 255)
    return fc_fail(ctx, FC_E_ARG, "bad scan parameters (asize=%d margin=%d maxdist=%d)", p->asize, p->margin, p->maxdist);
  if (b->n < 0 || b->n_words < 1 || b->max_l < 0) return fc_fail(ctx, FC_E_ARG, "bad batch (n=%lld n_words=%d max_l=%d)", (long long)b->n, b->n_words, b->max_l);
  if (b->max_l + 2 > FC_GENOME_PAD)
    return fc_fail(ctx, FC_E_ARG, "internal read length %d exceeds the genome padding (%d)", b->max_l, FC_GENOME_PAD);
  if (b->n > 0 && b->n_words * 32 < b->max_l) return fc_fail(ctx, FC_E_ARG, "n_words=%d too small for max_l=%d", b->n_words, b->max_l);
  if (b->n > 0 && (!b->d_meta || !b->d_reads)) return FC_E_ARG;
  return FC_OK;
}

extern "C" int fc_batch_words(int32_t max_l) { return batch_words(max_l); }

extern "C" int fc_batch_pack(fc_ctx* ctx, const fc_pairs* pr, const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b,
                             const uint8_t* d_frag, void* d_meta, uint32_t* d_reads, uint32_t* d_rn, uint32_t* d_q, void* stream) {
  if (!ctx || !pr || pr->n < 0 || !d_meta || !d_reads) return FC_E_ARG;
  if (!ctx->genome.loaded) return fc_fail(ctx, FC_E_NOGENOME, "fc_batch_pack before a genome was loaded (the descriptors hold genome coordinates)");
  if ((d_q != nullptr) != (d_q_a && d_q_b)) return FC_E_ARG;
  if (pr->n == 0) return FC_OK;
  const int nw = batch_words(pr->max_l > 0 ? pr->max_l : 0);
  fc::ReadView rv{pr->d_rlo, pr->d_rhi, pr->d_rn, pr->n, pr->n_words, pr->plane_stride > 0 ? pr->plane_stride : pr->n};
  pack_batch_kernel<<<(unsigned)((pr->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ctx->genome.view(), pr->n, pr->d_chrom, pr->d_a_start,
                                                                                    pr->d_b_end, pr->d_l, pr->d_flags, d_wden, d_frag, rv, nw,
                                                                                    (uint4*)d_meta, d_reads, d_rn, d_q_a, d_q_b, d_q);
  FC_LAUNCH_CHECK(ctx);
  return FC_OK;
}

extern "C" int fc_scan_batch(fc_ctx* ctx, const fc_scan_params* p, const fc_batch* b, fc_hit* d_out, void* stream) {
  int rc = check_batch(ctx, p, b);
  if (rc) return rc;
  if (b->n == 0) return FC_OK;
  BatchView v{(const uint4*)b->d_meta, b->d_reads, b->d_rn, b->n, b->n_words};
  return scan_launch(ctx, p, v, b->max_l, d_out, 0, Payload{}, (cudaStream_t)stream);
}

// scan + record a packed batch: into this context's aggregator, or -- connected to peers -- into the owners' buffers
static int scan_emit_batch(fc_ctx* ctx, const fc_scan_params* p, const fc_batch* b, fc_hit* d_out, const uint32_t* d_q,
                           const uint64_t* d_read_hash, const uint64_t* d_qname_hash, uint64_t idx_base, const uint64_t* d_idx,
                           cudaStream_t st) {
  int rc;
  BatchView v{(const uint4*)b->d_meta, b->d_reads, b->d_rn, b->n, b->n_words};
  if (ctx->agg.p2p_enabled) {  // connected to peers: every record goes to the rank that owns its key
    fc::P2PView pv;
    unsigned long long* overflow = nullptr;
    if ((rc = fc_agg_p2p_begin(ctx, &pv, &overflow))) return rc;
    if (b->n > 0) {
      Payload e{d_q, d_read_hash, d_qname_hash, idx_base, d_idx, nullptr, nullptr, overflow + 2};
      if ((rc = scan_launch(ctx, p, v, b->max_l, d_out, 2, e, st, &pv, overflow))) return rc;
    }
    fc_agg_p2p_end(ctx, idx_base, b->n);
    return FC_OK;
  }
  if (b->n == 0) return FC_OK;
  fc::EmitArgs ea{};
  if ((rc = fc_agg_emit_begin(ctx, b->n, st, &ea))) return rc;
  Payload e{d_q, d_read_hash, d_qname_hash, idx_base, d_idx, ea.n_recs, ea.recs, ea.n_recs + 6};
  if ((rc = scan_launch(ctx, p, v, b->max_l, d_out, 1, e, st))) return rc;
  fc_agg_emit_end(ctx, b->n, idx_base, d_idx != nullptr);
  return FC_OK;
}

extern "C" int fc_scan_emit_batch(fc_ctx* ctx, const fc_scan_params* p, const fc_batch* b, fc_hit* d_out, const uint32_t* d_q,
                                  const uint64_t* d_read_hash, const uint64_t* d_qname_hash, uint64_t idx_base, const uint64_t* d_idx,
                                  void* stream) {
  int rc = check_batch(ctx, p, b);
  if (rc) return rc;
  if (!d_out || !d_q || !d_read_hash) return FC_E_ARG;
  return scan_emit_batch(ctx, p, b, d_out, d_q, d_read_hash, d_qname_hash, idx_base, d_idx, (cudaStream_t)stream);
}

extern "C" int fc_scan_ties(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pr, const fc_hit* d_hits,
                            const int64_t* d_tie_off, fc_hit* d_ties, void* stream) {
  int rc = check_pairs(ctx, p, pr);
  if (rc) return rc;
  if (pr->n == 0) return FC_OK;
  fc::ScanCfg cfg{p->margin, p->maxdist, p->noncanonical, p->strandpref};
  int threads = 128;
  fc::ReadView rv{pr->d_rlo, pr->d_rhi, pr->d_rn, pr->n, pr->n_words, pr->plane_stride > 0 ? pr->plane_stride : pr->n};
  ties_kernel<<<(unsigned)((pr->n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      ctx->genome.view(), cfg, rv, pr->d_chrom, pr->d_a_start, pr->d_b_end, pr->d_l, pr->d_flags, d_hits, d_tie_off,
      d_ties);
  FC_LAUNCH_CHECK(ctx);
  return FC_OK;
}

// reference-facing call with HOST buffers (bench.py's e2e number goes through here)
extern "C" int fc_scan_host(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom,
                            const int32_t* h_a_start, const int32_t* h_b_end, const int32_t* h_l,
                            const uint8_t* h_flags, const uint8_t* h_ascii, int32_t stride, fc_hit* h_out) {
  if (!ctx || !p || n < 0) return FC_E_ARG;
  if (n == 0) return FC_OK;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  int32_t max_l = 0;
  for (int64_t i = 0; i < n; ++i) max_l = h_l[i] > max_l ? h_l[i] : max_l;
  if (max_l > stride) return fc_fail(ctx, FC_E_ARG, "stride %d smaller than the longest internal read %d", stride, max_l);
  int32_t n_words = (max_l + 31) / 32;
  if (n_words < 1) n_words = 1;
  size_t sizes[9] = {sizeof(int32_t) * (size_t)n, sizeof(int32_t) * (size_t)n, sizeof(int32_t) * (size_t)n,
                     sizeof(int32_t) * (size_t)n, (size_t)n + 4,           (size_t)n * (size_t)stride,
                     sizeof(uint32_t) * (size_t)n * n_words, sizeof(uint32_t) * (size_t)n * n_words,
                     sizeof(fc_hit) * (size_t)n};
  for (int k = 0; k < 9; ++k) FC_CUDA(ctx, ctx->host_path[k].reserve(sizes[k], st, false, 0));
  FC_CUDA(ctx, ctx->host_path[14].reserve(sizes[6], st, false, 0));
  int32_t* d_chrom = (int32_t*)ctx->host_path[0].p;
  int32_t* d_a = (int32_t*)ctx->host_path[1].p;
  int32_t* d_b = (int32_t*)ctx->host_path[2].p;
  int32_t* d_l = (int32_t*)ctx->host_path[3].p;
  uint8_t* d_fl = (uint8_t*)ctx->host_path[4].p;
  uint8_t* d_asc = (uint8_t*)ctx->host_path[5].p;
  uint32_t* d_rlo = (uint32_t*)ctx->host_path[6].p;
  uint32_t* d_rhi = (uint32_t*)ctx->host_path[7].p;
  uint32_t* d_rn = (uint32_t*)ctx->host_path[14].p;
  fc_hit* d_out = (fc_hit*)ctx->host_path[8].p;
  FC_CUDA(ctx, cudaMemcpyAsync(d_chrom, h_chrom, sizes[0], cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_a, h_a_start, sizes[1], cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_b, h_b_end, sizes[2], cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_l, h_l, sizes[3], cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_fl, h_flags, (size_t)n, cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(d_asc, h_ascii, sizes[5], cudaMemcpyHostToDevice, st));
  int rc = fc_pack_reads(ctx, n, d_asc, stride, d_l, n_words, d_rlo, d_rhi, d_rn, d_fl, st);
  if (rc) return rc;
  fc_pairs pr;
  pr.n = n;
  pr.d_chrom = d_chrom;
  pr.d_a_start = d_a;
  pr.d_b_end = d_b;
  pr.d_l = d_l;
  pr.d_flags = d_fl;
  pr.d_rlo = d_rlo;
  pr.d_rhi = d_rhi;
  pr.d_rn = d_rn;
  pr.n_words = n_words;
  pr.max_l = max_l;
  pr.plane_stride = 0;
  rc = fc_scan(ctx, p, &pr, d_out, st);
  if (rc) return rc;
  FC_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, sizes[8], cudaMemcpyDeviceToHost, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  return FC_OK;
}

// reference-facing batch call with HOST buffers: upload one batch of anchor pairs (scan inputs + aggregation payload),
// pack, scan, append the accepted junction records to the device aggregator, download the per-pair hits.
// This is what the python host's main loop calls per batch (the equivalent of find_circ.py:1295-1317, 1350-1378).
extern "C" int fc_batch_host(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom,
                             const int32_t* h_a_start, const int32_t* h_b_end, const int32_t* h_l, const uint8_t* h_flags,
                             const uint8_t* h_ascii, int32_t stride, const uint8_t* h_wden, const int16_t* h_q_a,
                             const int16_t* h_q_b, const uint64_t* h_read_hash, const uint64_t* h_qname_hash,
                             uint64_t idx_base, int32_t emit, fc_hit* h_out) {
  return fc_batch_host_idx(ctx, p, n, h_chrom, h_a_start, h_b_end, h_l, h_flags, h_ascii, stride, h_wden, h_q_a, h_q_b,
                           h_read_hash, h_qname_hash, nullptr, idx_base, emit, h_out);
}

extern "C" int fc_batch_host_idx(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom,
                                 const int32_t* h_a_start, const int32_t* h_b_end, const int32_t* h_l,
                                 const uint8_t* h_flags, const uint8_t* h_ascii, int32_t stride, const uint8_t* h_wden,
                                 const int16_t* h_q_a, const int16_t* h_q_b, const uint64_t* h_read_hash,
                                 const uint64_t* h_qname_hash, const uint64_t* h_idx, uint64_t idx_base, int32_t emit,
                                 fc_hit* h_out) {
  if (!ctx || !p || n < 0) return FC_E_ARG;
  ctx->last_n = 0;  // the retained batch is only valid after a call that succeeded
  if (n == 0) return FC_OK;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  int32_t max_l = 0;
  for (int64_t i = 0; i < n; ++i) max_l = h_l[i] > max_l ? h_l[i] : max_l;
  if (max_l > stride) return fc_fail(ctx, FC_E_ARG, "stride %d smaller than the longest internal read %d", stride, max_l);
  int32_t n_words = (max_l + 31) / 32;
  if (n_words < 1) n_words = 1;
  const size_t N = (size_t)n;
  size_t sizes[14] = {4 * N, 4 * N, 4 * N, 4 * N, N + 4, N * (size_t)stride, 4 * N * n_words, 4 * N * n_words,
                      sizeof(fc_hit) * N, N, 2 * N, 2 * N, 8 * N, 8 * N};
  for (int k = 0; k < 14; ++k) FC_CUDA(ctx, ctx->host_path[k].reserve(sizes[k], st, false, 0));
  FC_CUDA(ctx, ctx->host_path[14].reserve(sizes[6], st, false, 0));
  void* dp[14];
  for (int k = 0; k < 14; ++k) dp[k] = ctx->host_path[k].p;
  const void* src[14] = {h_chrom, h_a_start, h_b_end, h_l, h_flags, h_ascii, nullptr, nullptr, nullptr,
                         h_wden, h_q_a, h_q_b, h_read_hash, h_qname_hash};
  for (int k = 0; k < 14; ++k) {
    if (!src[k]) continue;
    if (k >= 9 && !src[k]) continue;
    size_t bytes = k == 4 ? N : sizes[k];
    FC_CUDA(ctx, cudaMemcpyAsync(dp[k], src[k], bytes, cudaMemcpyHostToDevice, st));
  }
  int rc = fc_pack_reads(ctx, n, (const uint8_t*)dp[5], stride, (const int32_t*)dp[3], n_words, (uint32_t*)dp[6],
                         (uint32_t*)dp[7], (uint32_t*)ctx->host_path[14].p, (uint8_t*)dp[4], st);
  if (rc) return rc;
  fc_pairs pr;
  pr.n = n;
  pr.d_chrom = (const int32_t*)dp[0];
  pr.d_a_start = (const int32_t*)dp[1];
  pr.d_b_end = (const int32_t*)dp[2];
  pr.d_l = (const int32_t*)dp[3];
  pr.d_flags = (const uint8_t*)dp[4];
  pr.d_rlo = (const uint32_t*)dp[6];
  pr.d_rhi = (const uint32_t*)dp[7];
  pr.d_rn = (const uint32_t*)ctx->host_path[14].p;
  pr.n_words = n_words;
  pr.max_l = max_l;
  pr.plane_stride = 0;
  rc = fc_scan(ctx, p, &pr, (fc_hit*)dp[8], st);
  if (rc) return rc;
  ctx->last_has_idx = false;
  if (h_idx) {
    FC_CUDA(ctx, ctx->host_path[15].reserve(8 * N + (size_t)n + 64, st, false, 0));
    FC_CUDA(ctx, cudaMemcpyAsync(ctx->host_path[15].p, h_idx, 8 * N, cudaMemcpyHostToDevice, st));
    ctx->last_has_idx = true;
  }
  if (emit) {
    if (!h_wden || !h_q_a || !h_q_b || !h_read_hash || !h_qname_hash) return fc_fail(ctx, FC_E_ARG, "emit needs the payload arrays");
    if (h_idx)
      rc = fc_agg_emit_idx(ctx, n, (const fc_hit*)dp[8], pr.d_chrom, pr.d_flags, (const uint8_t*)dp[9], (const int16_t*)dp[10],
                           (const int16_t*)dp[11], (const uint64_t*)dp[12], (const uint64_t*)dp[13], nullptr,
                           (const uint64_t*)ctx->host_path[15].p, st);
    else
      rc = fc_agg_emit(ctx, n, (const fc_hit*)dp[8], pr.d_chrom, pr.d_flags, (const uint8_t*)dp[9], (const int16_t*)dp[10],
                       (const int16_t*)dp[11], (const uint64_t*)dp[12], (const uint64_t*)dp[13], nullptr, idx_base, st);
    if (rc) return rc;
  }
  if (h_out) FC_CUDA(ctx, cudaMemcpyAsync(h_out, dp[8], sizes[8], cudaMemcpyDeviceToHost, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->last_n = n;
  ctx->last_pairs = pr;
  ctx->last_has_payload = h_wden && h_q_a && h_q_b && h_read_hash && h_qname_hash;
  return FC_OK;
}

// second half of a two-step batch: the host has looked at the hits of the last fc_batch_host(emit=0) call and decided
// per pair whether its first tie is recorded (the linear spans of a fragment are only recorded when its back-splices
// resolved to at most one junction, find_circ.py:1319-1329; --no-linear, :1328-1329)
extern "C" int fc_batch_emit_host(fc_ctx* ctx, const uint8_t* h_mask, uint64_t idx_base) {
  if (!ctx) return FC_E_ARG;
  if (ctx->last_n <= 0) return FC_OK;
  if (!ctx->last_has_payload) return fc_fail(ctx, FC_E_STATE, "the last batch was uploaded without aggregation payload");
  cudaStream_t st = ctx->own_stream;
  const int64_t n = ctx->last_n;
  uint8_t* d_mask = nullptr;
  const uint64_t* d_idx = ctx->last_has_idx ? (const uint64_t*)ctx->host_path[15].p : nullptr;
  if (h_mask) {
    // host_path[15] = [idx (8n bytes, optional) | mask (n bytes)]
    FC_CUDA(ctx, ctx->host_path[15].reserve(8 * (size_t)n + (size_t)n + 64, st, true, ctx->last_has_idx ? 8 * (size_t)n : 0));
    d_idx = ctx->last_has_idx ? (const uint64_t*)ctx->host_path[15].p : nullptr;
    d_mask = (uint8_t*)ctx->host_path[15].p + 8 * (size_t)n;
    FC_CUDA(ctx, cudaMemcpyAsync(d_mask, h_mask, (size_t)n, cudaMemcpyHostToDevice, st));
  }
  int rc;
  if (d_idx)
    rc = fc_agg_emit_idx(ctx, n, (const fc_hit*)ctx->host_path[8].p, ctx->last_pairs.d_chrom, ctx->last_pairs.d_flags,
                         (const uint8_t*)ctx->host_path[9].p, (const int16_t*)ctx->host_path[10].p,
                         (const int16_t*)ctx->host_path[11].p, (const uint64_t*)ctx->host_path[12].p,
                         (const uint64_t*)ctx->host_path[13].p, d_mask, d_idx, st);
  else
    rc = fc_agg_emit(ctx, n, (const fc_hit*)ctx->host_path[8].p, ctx->last_pairs.d_chrom, ctx->last_pairs.d_flags,
                     (const uint8_t*)ctx->host_path[9].p, (const int16_t*)ctx->host_path[10].p,
                     (const int16_t*)ctx->host_path[11].p, (const uint64_t*)ctx->host_path[12].p,
                     (const uint64_t*)ctx->host_path[13].p, d_mask, idx_base, st);
  if (rc) return rc;
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  return FC_OK;
}

// --all-hits on the last fc_batch_host batch: h_tie_off = exclusive prefix sum of n_hits (n+1 entries)
extern "C" int fc_batch_ties_host(fc_ctx* ctx, const fc_scan_params* p, const int64_t* h_tie_off, fc_hit* h_ties) {
  if (!ctx || !p || !h_tie_off || !h_ties) return FC_E_ARG;
  if (ctx->last_n <= 0) return FC_OK;
  cudaStream_t st = ctx->own_stream;
  const int64_t n = ctx->last_n;
  const int64_t total = h_tie_off[n];
  if (total <= 0) return FC_OK;
  // (the tie offsets have their own buffer: host_path[15] may hold the explicit stream positions of the batch)
  FC_CUDA(ctx, ctx->tie_off.reserve((size_t)(n + 1) * 8, st, false, 0));
  fc_dbuf& tb = ctx->agg.scratch[0];
  FC_CUDA(ctx, tb.reserve((size_t)total * sizeof(fc_hit), st, false, 0));
  FC_CUDA(ctx, cudaMemcpyAsync(ctx->tie_off.p, h_tie_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
  int rc = fc_scan_ties(ctx, p, &ctx->last_pairs, (const fc_hit*)ctx->host_path[8].p, (const int64_t*)ctx->tie_off.p,
                        (fc_hit*)tb.p, st);
  if (rc) return rc;
  FC_CUDA(ctx, cudaMemcpyAsync(h_ties, tb.p, (size_t)total * sizeof(fc_hit), cudaMemcpyDeviceToHost, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  return FC_OK;
}

// rows that arrive with their internal read part already packed (native ingest): no ASCII upload, no pack kernel
extern "C" int fc_batch_host_planes(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom,
                                    const int32_t* h_a_start, const int32_t* h_b_end, const int32_t* h_l,
                                    const uint8_t* h_flags, const uint32_t* h_rlo, const uint32_t* h_rhi,
                                    const uint32_t* h_rn, int32_t n_words, int64_t plane_stride, int32_t max_l,
                                    const uint8_t* h_wden, const int16_t* h_q_a, const int16_t* h_q_b,
                                    const uint64_t* h_read_hash, const uint64_t* h_qname_hash, const uint64_t* h_idx,
                                    uint64_t idx_base, int32_t emit, fc_hit* h_out) {
  if (!ctx || !p || n < 0 || n_words < 1 || plane_stride < n) return FC_E_ARG;
  ctx->last_n = 0;  // the retained batch is only valid after a call that succeeded
  if (n == 0) return FC_OK;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  const size_t N = (size_t)n;
  size_t sizes[14] = {4 * N, 4 * N, 4 * N, 4 * N, N + 4, 16, 4 * N * n_words, 4 * N * n_words,
                      sizeof(fc_hit) * N, N, 2 * N, 2 * N, 8 * N, 8 * N};
  for (int k = 0; k < 14; ++k) FC_CUDA(ctx, ctx->host_path[k].reserve(sizes[k], st, false, 0));
  FC_CUDA(ctx, ctx->host_path[14].reserve(sizes[6], st, false, 0));
  void* dp[15];
  for (int k = 0; k < 15; ++k) dp[k] = ctx->host_path[k].p;
  const bool payload = h_wden && h_q_a && h_q_b && h_read_hash && h_qname_hash;
  if (emit && !payload) return fc_fail(ctx, FC_E_ARG, "emit needs the payload arrays");
  uint64_t* d_idx = nullptr;
  if (h_idx) {
    FC_CUDA(ctx, ctx->host_path[15].reserve(8 * N, st, false, 0));
    d_idx = (uint64_t*)ctx->host_path[15].p;
  }
  int rc = FC_OK;
  if (emit && (rc = fc_agg_reserve_records(ctx, n, st))) return rc;  // no growth while chunks are in flight
  // The batch goes through in chunks on two streams: while chunk k is scanned and its hits travel back, the columns of
  // chunk k+1 are already on their way in (PCIe is full duplex and the copy engines run beside the kernels).
  if (!ctx->own_stream2) {
    FC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->own_stream2, cudaStreamNonBlocking));
    FC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk[0], cudaEventDisableTiming));
    FC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk[1], cudaEventDisableTiming));
  }
  FC_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[0], st));  // everything queued on the main stream so far comes first
  FC_CUDA(ctx, cudaStreamWaitEvent(ctx->own_stream2, ctx->ev_chunk[0], 0));
  // pairs per chunk (FC_HOST_CHUNK overrides): copies of 4 MiB and more per column run at the full PCIe rate, smaller ones do
  // not, and that outweighs what the overlap of copies and kernels buys
  const int64_t chunk = ctx->host_chunk;  // (fc_ctx_create: FC_HOST_CHUNK or 1 M pairs -- measured on the B200 box at 1 M pairs: 256k 1.63 ms, 512k 1.56 ms, one chunk 1.51 ms)
  const size_t elem[14] = {4, 4, 4, 4, 1, 0, 0, 0, 0, 1, 2, 2, 8, 8};
  const void* src[14] = {h_chrom, h_a_start, h_b_end, h_l, h_flags, nullptr, nullptr, nullptr, nullptr,
                         h_wden, h_q_a, h_q_b, h_read_hash, h_qname_hash};
  const uint32_t* hp[3] = {h_rlo, h_rhi, h_rn};
  void* dpl[3] = {dp[6], dp[7], dp[14]};
  int which = 0;
  for (int64_t c0 = 0; c0 < n; c0 += chunk, which ^= 1) {
    const int64_t cn = n - c0 < chunk ? n - c0 : chunk;
    cudaStream_t cs = which ? ctx->own_stream2 : st;
    for (int k = 0; k < 14; ++k) {
      if (!src[k]) continue;
      FC_CUDA(ctx, cudaMemcpyAsync((char*)dp[k] + elem[k] * c0, (const char*)src[k] + elem[k] * c0, elem[k] * cn,
                                   cudaMemcpyHostToDevice, cs));
    }
    // planes: word rows of n entries each (the host arrays may be strided wider)
    for (int k = 0; k < 3; ++k)
      FC_CUDA(ctx, cudaMemcpy2DAsync((char*)dpl[k] + 4 * c0, 4 * N, (const char*)hp[k] + 4 * c0, 4 * (size_t)plane_stride, 4 * cn,
                                     (size_t)n_words, cudaMemcpyHostToDevice, cs));
    if (d_idx) FC_CUDA(ctx, cudaMemcpyAsync(d_idx + c0, h_idx + c0, 8 * cn, cudaMemcpyHostToDevice, cs));
    fc_pairs pc;
    pc.n = cn;
    pc.d_chrom = (const int32_t*)dp[0] + c0;
    pc.d_a_start = (const int32_t*)dp[1] + c0;
    pc.d_b_end = (const int32_t*)dp[2] + c0;
    pc.d_l = (const int32_t*)dp[3] + c0;
    pc.d_flags = (const uint8_t*)dp[4] + c0;
    pc.d_rlo = (const uint32_t*)dp[6] + c0;
    pc.d_rhi = (const uint32_t*)dp[7] + c0;
    pc.d_rn = (const uint32_t*)dp[14] + c0;
    pc.n_words = n_words;
    pc.max_l = max_l;
    pc.plane_stride = n;
    fc_hit* d_hits = (fc_hit*)dp[8] + c0;
    if (emit) {  // scan and record in one kernel
      rc = scan_emit_soa(ctx, p, &pc, d_hits, (const uint8_t*)dp[9] + c0, (const int16_t*)dp[10] + c0, (const int16_t*)dp[11] + c0,
                         (const uint64_t*)dp[12] + c0, (const uint64_t*)dp[13] + c0, idx_base + (uint64_t)c0,
                         d_idx ? d_idx + c0 : nullptr, cs, c0, n);
    } else {
      BatchView bv;
      rc = pack_soa(ctx, &pc, nullptr, nullptr, nullptr, nullptr, cs, c0, n, &bv, nullptr);
      if (!rc) rc = scan_launch(ctx, p, bv, max_l, d_hits, 0, Payload{}, cs);
    }
    if (rc) {
      cudaStreamSynchronize(ctx->own_stream2);  // nothing of this call stays in flight behind the error
      cudaStreamSynchronize(st);
      return rc;
    }
    if (h_out) FC_CUDA(ctx, cudaMemcpyAsync(h_out + c0, d_hits, sizeof(fc_hit) * cn, cudaMemcpyDeviceToHost, cs));
  }
  FC_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[1], ctx->own_stream2));
  FC_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_chunk[1], 0));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  fc_pairs pr;
  pr.n = n;
  pr.d_chrom = (const int32_t*)dp[0];
  pr.d_a_start = (const int32_t*)dp[1];
  pr.d_b_end = (const int32_t*)dp[2];
  pr.d_l = (const int32_t*)dp[3];
  pr.d_flags = (const uint8_t*)dp[4];
  pr.d_rlo = (const uint32_t*)dp[6];
  pr.d_rhi = (const uint32_t*)dp[7];
  pr.d_rn = (const uint32_t*)dp[14];
  pr.n_words = n_words;
  pr.max_l = max_l;
  pr.plane_stride = 0;
  ctx->last_n = n;
  ctx->last_pairs = pr;
  ctx->last_has_payload = payload;
  ctx->last_has_idx = h_idx != nullptr;
  return FC_OK;
}

// ---------------------------------------------------------------- streamed host batches (fc_stream)
// The reference's main loop reads, scans and records one fragment after the other (find_circ.py:1535-1574); here the host
// fills batch k+1 while batch k is copied, scanned, recorded and its hits travel back -- every slot has its own stream.
struct fc_stream_slot {
  cudaStream_t st = nullptr;
  cudaEvent_t done = nullptr;
  bool busy = false;
  fc_dbuf buf[12];  // meta, reads, q, read_hash, qname_hash, idx, rn (dense), rn_idx, rn_rows, hits, compact, masks
};
struct fc_stream {
  fc_ctx* ctx = nullptr;
  int64_t cap = 0;
  int32_t max_words = 0;
  std::vector<fc_stream_slot> slots;
};

namespace {
__global__ void scatter_rn_kernel(int64_t n_rn, const uint32_t* __restrict__ rows_idx, const uint32_t* __restrict__ rows, int nw,
                                  uint32_t* __restrict__ dense) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rn * nw) return;
  const int64_t r = t / nw;
  const int k = (int)(t - r * nw);
  dense[(int64_t)rows_idx[r] * nw + k] = rows[t];
}

// fc_hit[n] -> (start, end) pairs + one bit per pair "has a breakpoint" + one bit per pair "minus strand"
__global__ void compact_hits_kernel(int64_t n, const fc_hit* __restrict__ hits, int2* __restrict__ se, uint32_t* __restrict__ hit_mask,
                                    uint32_t* __restrict__ strand_mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool hit = false, minus = false;
  if (i < n) {
    const uint4 h = reinterpret_cast<const uint4*>(hits)[i];
    hit = (h.z & 0xFFFFu) != 0u;
    minus = hit && (h.w & 1u);
    se[i] = make_int2((int)h.x, (int)h.y);
  }
  const unsigned hm = __ballot_sync(0xffffffffu, hit), sm = __ballot_sync(0xffffffffu, minus);
  if ((threadIdx.x & 31) == 0 && i < n) {
    hit_mask[i >> 5] = hm;
    strand_mask[i >> 5] = sm;
  }
}
}  // namespace

extern "C" int fc_stream_create(fc_ctx* ctx, int32_t n_slots, int64_t cap_rows, int32_t max_words, fc_stream** out) {
  if (!ctx || !out || n_slots < 1 || n_slots > 16 || cap_rows < 1 || max_words < 1) return FC_E_ARG;
  *out = nullptr;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = fc_agg_reserve_records(ctx, 0, ctx->own_stream);  // the counters exist before any slot's stream touches them
  if (rc) return rc;
  FC_CUDA(ctx, cudaStreamSynchronize(ctx->own_stream));
  fc_stream* fs = new fc_stream();
  fs->ctx = ctx;
  fs->cap = cap_rows;
  fs->max_words = max_words;
  fs->slots.resize(n_slots);
  const size_t C = (size_t)cap_rows, W = (size_t)max_words;
  const size_t sizes[12] = {16 * C, 8 * C * W, 4 * C, 8 * C, 8 * C, 8 * C, 4 * C * W, 4 * C, 4 * C * W, 16 * C, 8 * C, 8 * ((C + 31) / 32) + 64};
  for (auto& sl : fs->slots) {
    cudaError_t e = cudaStreamCreateWithFlags(&sl.st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming);
    for (int k = 0; k < 12 && e == cudaSuccess; ++k) e = sl.buf[k].reserve(sizes[k], sl.st, false, 0);
    if (e != cudaSuccess) {
      fc_stream_destroy(fs);
      return fc_fail(ctx, FC_E_CUDA, "fc_stream_create: %s", cudaGetErrorString(e));
    }
  }
  *out = fs;
  return FC_OK;
}

extern "C" void fc_stream_destroy(fc_stream* fs) {
  if (!fs) return;
  cudaSetDevice(fs->ctx->device);
  for (auto& sl : fs->slots) {
    if (sl.st) cudaStreamSynchronize(sl.st);
    for (auto& b : sl.buf) b.release();
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.st) cudaStreamDestroy(sl.st);
  }
  delete fs;
}

extern "C" int fc_stream_submit(fc_stream* fs, int32_t slot, const fc_scan_params* p, const fc_host_batch* hb) {
  if (!fs || !hb || !p || slot < 0 || slot >= (int32_t)fs->slots.size()) return FC_E_ARG;
  fc_ctx* ctx = fs->ctx;
  fc_stream_slot& sl = fs->slots[slot];
  if (sl.busy) return fc_fail(ctx, FC_E_STATE, "fc_stream_submit: slot %d is in flight (fc_stream_wait first)", slot);
  const int64_t n = hb->n;
  if (n < 0 || n > fs->cap || hb->n_words < 1 || hb->n_words > fs->max_words || hb->n_rn < 0 || hb->n_rn > n)
    return fc_fail(ctx, FC_E_ARG, "fc_stream_submit: batch of %lld rows x %d words does not fit the slots (%lld x %d)", (long long)n, hb->n_words,
                   (long long)fs->cap, fs->max_words);
  if (n == 0) return FC_OK;
  if (!hb->meta || !hb->reads || (hb->emit && (!hb->q || !hb->read_hash)) || (hb->n_rn && (!hb->rn_idx || !hb->rn_rows))) return FC_E_ARG;
  if (hb->out_mode == 1 && !hb->out_hits) return FC_E_ARG;
  if (hb->out_mode == 2 && (!hb->out_hits || !hb->out_hit_mask || !hb->out_strand_mask)) return FC_E_ARG;
  FC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = sl.st;
  const size_t N = (size_t)n, W = (size_t)hb->n_words;
  auto dp = [&](int k) { return sl.buf[k].p; };
  FC_CUDA(ctx, cudaMemcpyAsync(dp(0), hb->meta, 16 * N, cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaMemcpyAsync(dp(1), hb->reads, 8 * N * W, cudaMemcpyHostToDevice, st));
  if (hb->emit) {
    FC_CUDA(ctx, cudaMemcpyAsync(dp(2), hb->q, 4 * N, cudaMemcpyHostToDevice, st));
    FC_CUDA(ctx, cudaMemcpyAsync(dp(3), hb->read_hash, 8 * N, cudaMemcpyHostToDevice, st));
    if (hb->qname_hash) FC_CUDA(ctx, cudaMemcpyAsync(dp(4), hb->qname_hash, 8 * N, cudaMemcpyHostToDevice, st));
    if (hb->idx) FC_CUDA(ctx, cudaMemcpyAsync(dp(5), hb->idx, 8 * N, cudaMemcpyHostToDevice, st));
  }
  if (hb->n_rn) {  // the N planes of the few reads that have any: a sparse list, scattered into the dense rows the scan reads
    FC_CUDA(ctx, cudaMemcpyAsync(dp(7), hb->rn_idx, 4 * (size_t)hb->n_rn, cudaMemcpyHostToDevice, st));
    FC_CUDA(ctx, cudaMemcpyAsync(dp(8), hb->rn_rows, 4 * (size_t)hb->n_rn * W, cudaMemcpyHostToDevice, st));
    const int64_t t = hb->n_rn * (int64_t)W;
    scatter_rn_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(hb->n_rn, (const uint32_t*)dp(7), (const uint32_t*)dp(8), (int)W,
                                                                 (uint32_t*)dp(6));
    FC_LAUNCH_CHECK(ctx);
  }
  fc_batch b{n, dp(0), (const uint32_t*)dp(1), (const uint32_t*)dp(6), hb->n_words, hb->max_l};
  int rc = check_batch(ctx, p, &b);
  if (rc) return rc;
  fc_hit* d_hits = (fc_hit*)dp(9);
  if (hb->emit) {
    fc_agg& a = ctx->agg;
    if (!a.p2p_enabled && (size_t)(a.n_recs + n) * sizeof(fc_jrec) > a.recs.cap) {
      // the record buffer has to grow: nothing may be in flight on the other slots' streams while it moves
      FC_CUDA(ctx, cudaDeviceSynchronize());
      const int64_t extra = a.n_recs + n > 2 * n ? a.n_recs + n : 2 * n;  // (at least doubling)
      if ((rc = fc_agg_reserve_records(ctx, extra, st))) return rc;
    }
    rc = scan_emit_batch(ctx, p, &b, d_hits, (const uint32_t*)dp(2), (const uint64_t*)dp(3), hb->qname_hash ? (const uint64_t*)dp(4) : nullptr,
                         hb->idx_base, hb->idx ? (const uint64_t*)dp(5) : nullptr, st);
  } else {
    BatchView v{(const uint4*)b.d_meta, b.d_reads, b.d_rn, b.n, b.n_words};
    rc = scan_launch(ctx, p, v, b.max_l, d_hits, 0, Payload{}, st);
  }
  if (rc) return rc;
  if (hb->out_mode == 1) {
    FC_CUDA(ctx, cudaMemcpyAsync(hb->out_hits, d_hits, 16 * N, cudaMemcpyDeviceToHost, st));
  } else if (hb->out_mode == 2) {
    uint32_t* masks = (uint32_t*)dp(11);
    const size_t mw = (N + 31) / 32;
    compact_hits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d_hits, (int2*)dp(10), masks, masks + mw);
    FC_LAUNCH_CHECK(ctx);
    FC_CUDA(ctx, cudaMemcpyAsync(hb->out_hits, dp(10), 8 * N, cudaMemcpyDeviceToHost, st));
    FC_CUDA(ctx, cudaMemcpyAsync(hb->out_hit_mask, masks, 4 * mw, cudaMemcpyDeviceToHost, st));
    FC_CUDA(ctx, cudaMemcpyAsync(hb->out_strand_mask, masks + mw, 4 * mw, cudaMemcpyDeviceToHost, st));
  }
  FC_CUDA(ctx, cudaEventRecord(sl.done, st));
  sl.busy = true;
  return FC_OK;
}

extern "C" int fc_stream_wait(fc_stream* fs, int32_t slot) {
  if (!fs || slot < 0 || slot >= (int32_t)fs->slots.size()) return FC_E_ARG;
  fc_stream_slot& sl = fs->slots[slot];
  if (!sl.busy) return FC_OK;
  FC_CUDA(fs->ctx, cudaEventSynchronize(sl.done));
  sl.busy = false;
  return FC_OK;
}

extern "C" int fc_stream_query(fc_stream* fs, int32_t slot) {
  if (!fs || slot < 0 || slot >= (int32_t)fs->slots.size()) return FC_E_ARG;
  fc_stream_slot& sl = fs->slots[slot];
  if (!sl.busy) return 1;
  const cudaError_t e = cudaEventQuery(sl.done);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) return 0;
  return fc_fail(fs->ctx, FC_E_CUDA, "fc_stream_query: %s", cudaGetErrorString(e));
}
