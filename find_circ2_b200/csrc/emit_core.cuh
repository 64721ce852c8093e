// emit_core.cuh -- turning the accepted pairs of one CTA into 48-byte junction records (fc_jrec), shared by the stand-alone
// emit kernel (agg.cu) and by the scan kernel that emits on the way (scan.cu, fc_scan_emit).
//
// Replaces the argument marshalling of SpliceSiteStorage.add / Hit.add (/root/reference/find_circ.py:526-582, 681-690).
#pragma once
#include <stdint.h>

#include "../../include/findcirc_b200.h"

namespace fc {

struct EmitArgs {
  const uint8_t* wden;         // weight denominator (find_circ.py:1084)
  const int16_t* q_a;          // AS - XS of anchor A / B (find_circ.py:556-559)
  const int16_t* q_b;
  const uint64_t* read_hash;
  const uint64_t* qname_hash;
  uint64_t idx_base;           // idx = idx_base + pair index, unless `idx` gives it explicitly
  const uint64_t* idx;
  unsigned long long* n_recs;  // record counter of the context (device)
  fc_jrec* recs;               // record buffer of the context
};

// Called by EVERY thread of a CTA of BS threads (BS a multiple of 32, at most 1024).  `accept` threads hand over their
// pair (index i, hit words, chromosome id, pair flags).  The CTA claims its slots with one atomic on the record counter,
// groups the records in shared memory and writes them as one run of consecutive 16-byte stores; the buffer is therefore
// NOT in stream order -- every consumer orders by fc_jrec.idx where order matters.
template <int BS>
__device__ __forceinline__ void emit_block(bool accept, int64_t i, int32_t h_start, int32_t h_end, uint32_t w2, uint32_t w3,
                                           uint32_t chrom, uint32_t pair_flags, const EmitArgs& e) {
  __shared__ unsigned int s_warp[BS / 32];
  __shared__ unsigned int s_total;
  __shared__ unsigned long long s_base;
  __shared__ uint4 s_rec[BS * 3];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned ballot = __ballot_sync(0xffffffffu, accept);
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int total = 0;
#pragma unroll
    for (int w = 0; w < BS / 32; ++w) {
      const unsigned int c = s_warp[w];
      s_warp[w] = total;
      total += c;
    }
    s_total = total;
    s_base = total ? atomicAdd(e.n_recs, (unsigned long long)total) : 0ull;
  }
  __syncthreads();
  if (accept) {
    const bool backsplice = pair_flags & FC_PF_BACKSPLICE;
    fc_jrec r;
    r.chrom = chrom;
    r.start = (uint32_t)h_start;
    r.end = (uint32_t)h_end;
    const uint32_t strand = w3 & 1u, sig = (w3 >> 1) & 0xFFFu;
    const uint64_t rh = e.read_hash[i];
    r.sk = strand | (backsplice ? 0u : 2u) | ((uint32_t)(rh & 1ull) << 2) | ((uint32_t)e.wden[i] << 8) | (sig << 16);
    r.idx = e.idx ? e.idx[i] : e.idx_base + (uint64_t)i;
    r.read_hash = rh;
    r.qname_hash = e.qname_hash[i];
    // by convention A precedes B in the genome: swap for back-splices (find_circ.py:552-553)
    r.q_left = backsplice ? e.q_b[i] : e.q_a[i];
    r.q_right = backsplice ? e.q_a[i] : e.q_b[i];
    r.n_hits = (uint16_t)(w2 & 0xFFFFu);
    r.dist = (uint8_t)((w2 >> 16) & 0xFFu);
    r.ov = (uint8_t)(w2 >> 24);
    uint4* stage = s_rec + (size_t)(s_warp[warp] + __popc(ballot & ((1u << lane) - 1u))) * 3;
    const uint4* src = reinterpret_cast<const uint4*>(&r);
    stage[0] = src[0];
    stage[1] = src[1];
    stage[2] = src[2];
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(e.recs + s_base);
  for (unsigned int w = threadIdx.x; w < s_total * 3u; w += BS) out[w] = s_rec[w];
}

}  // namespace fc
