// emit_core.cuh -- turning the accepted pairs of one CTA into 48-byte junction records (fc_jrec), shared by the stand-alone
// emit kernel (agg.cu) and by the scan kernel that emits on the way (scan.cu, fc_scan_emit).
//
// Replaces the argument marshalling of SpliceSiteStorage.add / Hit.add (/root/reference/find_circ.py:526-582, 681-690).
#pragma once
#include <stdint.h>

#include "../../include/findcirc_b200.h"

// 64-bit mix (splitmix64 finaliser)
__host__ __device__ inline uint64_t fc_mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}
__host__ __device__ inline uint64_t fc_key_hash(uint32_t chrom, uint32_t start, uint32_t end, uint32_t sk, uint64_t seed) {
  uint64_t a = ((uint64_t)chrom << 32) | start;
  uint64_t b = ((uint64_t)end << 32) | (sk & 3u);
  return fc_mix64(fc_mix64(a + seed) ^ (b * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL));
}

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

// ---- bulk asynchronous copies shared memory -> global memory (TMA, UBLKCP): one instruction moves a CTA's whole run of
// records, to local HBM or to a peer over NVLink, instead of hundreds of 16-byte stores through the LSU
__device__ __forceinline__ void smem_writes_before_bulk() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {  // 16-byte aligned, bytes % 16 == 0
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_and_release() {  // the issuing thread may leave once the source has been read
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Called by EVERY thread of a CTA of BS threads (BS a multiple of 32, at most 1024).  `accept` threads hand over their
// record.  The CTA claims its slots with one atomic on the record counter, groups the records in shared memory and
// writes them as ONE bulk copy; the buffer is therefore NOT in stream order -- every consumer orders by fc_jrec.idx
// where order matters.
template <int BS>
__device__ __forceinline__ void emit_block(bool accept, const fc_jrec& r, unsigned long long* n_recs, fc_jrec* recs) {
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
    s_base = total ? atomicAdd(n_recs, (unsigned long long)total) : 0ull;
  }
  __syncthreads();
  if (accept) {
    uint4* stage = s_rec + (size_t)(s_warp[warp] + __popc(ballot & ((1u << lane) - 1u))) * 3;
    const uint4* src = reinterpret_cast<const uint4*>(&r);
    stage[0] = src[0];
    stage[1] = src[1];
    stage[2] = src[2];
    smem_writes_before_bulk();
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_total) {
    bulk_store(recs + s_base, s_rec, s_total * (unsigned int)sizeof(fc_jrec));
    bulk_commit_and_release();
  }
}

// ---- the same towards the rank that owns the junction key (fused emit + exchange over peer memory) --------------------
// Every rank's record buffer is cut into 2 x world SLICES of slice_cap records: slice (parity, s) of rank d receives the
// records that source rank s sends to owner d in the steps of that parity.  A source therefore allocates its slots with
// counters in its OWN memory (device-scope atomics, no NVLink round trip) and only the 48-byte records cross the wire;
// the counts are published to the owners by the barrier kernel that ends the step (agg.cu: p2p_barrier_kernel).  Two
// parities: a rank may already write step k+1 into a peer that still reduces step k.
struct P2PView {
  fc_jrec* recs[8];             // record buffer of every rank (peer memory)
  unsigned long long* cnt[8];   // counter block of every rank (peer memory)
  unsigned long long slice_cap; // records per slice
  int world;
  int rank;
  int parity;                   // 0 / 1: which half of the buffers this step uses
};
constexpr int FC_CNT_WORDS = 256;      // 64-bit words of a context's counter block
constexpr int FC_CNT_SRC = 64;         // word FC_CNT_SRC + FC_CNT_SRC_STRIDE * d: records this rank has sent to destination d in this step;
constexpr int FC_CNT_SRC_STRIDE = 16;  // one 128-byte line each: every CTA of the scan bumps all of them, a shared sector would serialise in one L2 slice
constexpr int FC_CNT_SLICE = 40;  // counter words [40, 56): [parity][source] records received from source in the step

// one Hit.add() call as a record (find_circ.py:526-582): hit words of the scan + the payload of the pair
__device__ __forceinline__ fc_jrec make_record_from(int32_t h_start, int32_t h_end, uint32_t w2, uint32_t w3, uint32_t chrom,
                                                    uint32_t pair_flags, uint32_t wden, int16_t q_a, int16_t q_b, uint64_t rh,
                                                    uint64_t qh, uint64_t idx) {
  const bool backsplice = pair_flags & FC_PF_BACKSPLICE;
  fc_jrec r;
  r.chrom = chrom;
  r.start = (uint32_t)h_start;
  r.end = (uint32_t)h_end;
  const uint32_t strand = w3 & 1u, sig = (w3 >> 1) & 0xFFFu;
  r.sk = strand | (backsplice ? 0u : 2u) | ((uint32_t)(rh & 1ull) << 2) | (wden << 8) | (sig << 16);
  r.idx = idx;
  r.read_hash = rh;
  r.qname_hash = qh;
  // by convention A precedes B in the genome: swap for back-splices (find_circ.py:552-553)
  r.q_left = backsplice ? q_b : q_a;
  r.q_right = backsplice ? q_a : q_b;
  r.n_hits = (uint16_t)(w2 & 0xFFFFu);
  r.dist = (uint8_t)((w2 >> 16) & 0xFFu);
  r.ov = (uint8_t)(w2 >> 24);
  return r;
}
__device__ __forceinline__ fc_jrec make_record(int64_t i, int32_t h_start, int32_t h_end, uint32_t w2, uint32_t w3, uint32_t chrom,
                                               uint32_t pair_flags, const EmitArgs& e) {
  return make_record_from(h_start, h_end, w2, w3, chrom, pair_flags, e.wden[i], e.q_a[i], e.q_b[i], e.read_hash[i], e.qname_hash[i],
                          e.idx ? e.idx[i] : e.idx_base + (uint64_t)i);
}

// Called by every thread of a CTA of BS threads.  The CTA's records are grouped by destination rank in shared memory;
// one device-scope atomic per CTA and destination on the SOURCE's own counters claims the slots of the source's slice
// in the owner's buffer; every group then goes out as one run of consecutive 16-byte stores -- full-size write packets
// on NVLink instead of scattered 16-byte ones, and nothing on the CTA's critical path crosses the wire.
template <int BS>
__device__ __forceinline__ void emit_p2p_block(bool accept, const fc_jrec& r, const P2PView& pv, unsigned long long* overflow) {
  __shared__ uint4 s_rec[BS * 3];
  __shared__ unsigned int s_cnt[8], s_off[9];
  __shared__ unsigned long long s_base[8];
  int dest = 0;
  if (accept) dest = (int)(fc_key_hash(r.chrom, r.start, r.end, r.sk, 0x5bd1e995ULL) % (uint64_t)pv.world);
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  unsigned int local = 0;
  {
    // warp-aggregated shared-memory allocation: rank of this record among the CTA's records for its destination
    const unsigned lane = threadIdx.x & 31;
    const unsigned amask = __ballot_sync(0xffffffffu, accept);
    if (accept) {
      const unsigned peers = __match_any_sync(amask, dest);
      const int leader = __ffs((int)peers) - 1;
      unsigned int wbase = 0;
      if ((int)lane == leader) wbase = atomicAdd(&s_cnt[dest], (unsigned int)__popc(peers));
      wbase = __shfl_sync(peers, wbase, leader);
      local = wbase + (unsigned int)__popc(peers & ((1u << lane) - 1u));
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < pv.world && s_cnt[threadIdx.x])
    s_base[threadIdx.x] = atomicAdd(pv.cnt[pv.rank] + FC_CNT_SRC + FC_CNT_SRC_STRIDE * threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
  if (threadIdx.x == 0) {
    unsigned int acc = 0;
    for (int d = 0; d < 8; ++d) {
      s_off[d] = acc;
      acc += d < pv.world ? s_cnt[d] : 0u;
    }
    s_off[8] = acc;
  }
  __syncthreads();
  if (accept) {
    const uint4* src = reinterpret_cast<const uint4*>(&r);
    uint4* dst = s_rec + (size_t)(s_off[dest] + local) * 3;
    dst[0] = src[0];
    dst[1] = src[1];
    dst[2] = src[2];
    smem_writes_before_bulk();
  }
  __syncthreads();
  // one bulk copy per destination: the run goes over NVLink (or into local HBM) in full-size packets and no thread of
  // the CTA spends load/store slots on it
  if ((int)threadIdx.x < pv.world && s_cnt[threadIdx.x]) {
    const int d = (int)threadIdx.x;
    const unsigned long long base = s_base[d];
    unsigned long long room = base < pv.slice_cap ? pv.slice_cap - base : 0ull;
    const unsigned int cnt = s_cnt[d] < room ? s_cnt[d] : (unsigned int)room;
    if (cnt < s_cnt[d]) atomicAdd(overflow, (unsigned long long)(s_cnt[d] - cnt));
    if (cnt) {
      const unsigned long long slice0 = (unsigned long long)(pv.parity * pv.world + pv.rank) * pv.slice_cap;
      bulk_store(pv.recs[d] + slice0 + base, s_rec + (size_t)s_off[d] * 3, cnt * (unsigned int)sizeof(fc_jrec));
      bulk_commit_and_release();
    }
  }
}

}  // namespace fc
