// agg.cu -- junction aggregation on the device.
//
// Replaces SpliceSiteStorage.add / Hit.add and the Hit reductions (/root/reference/find_circ.py:486-600, 657-690):
// instead of one Python dict insert + list appends per accepted span, accepted spans become 48-byte records
// (fc_jrec) and fc_agg_finalize() reduces them per junction (chrom,start,end,strand,kind):
//     n_spanned            count                                                   (:543)
//     n_weighted           sum of weights, exact: weights 1/den with den a power of two are order independent;
//                          inputs holding any other denominator are re-summed sequentially in stream order (:544)
//     n_uniq_bridges       same, over records with both anchor qualities non-zero  (:561-563)
//     best_qual_left/right max                                                     (:592-593)
//     edits/overlap/n_hits min                                                     (:727)
//     first idx            min  -> discovery order -> junction name                (:684-686)
//     n_frags              distinct qname hashes                                   (:584-586)
//     n_uniq               distinct strand-invariant read hashes (palindromes count half, :588-590)
// Two implementations with identical results: a sort-free one (one pass: 128-bit CAS hash tables, per-junction
// accumulators updated with integer atomics, default) and a sort-based one (stable CUB radix sort + warp-segmented
// reduction + sequential float replay, fallback).  Records are produced by emit_core.cuh -- by a stand-alone kernel here
// or inside the scan kernel (scan.cu); multi-GPU: straight into the owner rank's buffer over NVLink.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <cub/cub.cuh>

#include "fc_internal.cuh"

namespace {

struct JAcc {  // per-junction accumulators of the sort-based path, filled with integer atomics (deterministic)
  unsigned int n_spanned;
  unsigned int cw[4];  // records with weight denominator 1,2,4,8
  unsigned int cw_other;
  unsigned int cb[4];  // same, restricted to unique bridges
  unsigned int cb_other;
  int max_ql, max_qr;
  unsigned int min_dist, min_ov, min_nh;
  unsigned int n_frags, n_uniq, n_pal;
  unsigned long long first_idx;
  unsigned int seg_start;  // index of the first sorted record
  unsigned int pad;
};

// ---- 128-bit table entries and compare-and-swap (ATOMG.E.CAS.128) ------------------------------------------------
struct alignas(16) U128 {
  unsigned long long lo, hi;
};
__device__ __forceinline__ U128 cas128(U128* addr, U128 cmp, U128 val) {
  U128 old;
  asm volatile(
      "{\n\t"
      ".reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t"
      "}\n"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr)
      : "memory");
  return old;
}

// One record per accepted pair (the scan found a breakpoint and the caller's mask, if any, keeps the pair); see
// emit_core.cuh for how a CTA writes its records.
__global__ void __launch_bounds__(256) emit_kernel(int64_t n, const fc_hit* __restrict__ hits, const uint8_t* __restrict__ mask,
                                                   const int32_t* __restrict__ chrom, const uint8_t* __restrict__ flags,
                                                   fc::EmitArgs e) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  fc_hit h = {0, 0, 0u, 0u};
  bool accept = false;
  uint32_t c = 0, fl = 0;
  if (i < n) {
    h = hits[i];
    accept = (h.w2 & 0xFFFFu) != 0u && (!mask || mask[i]);
    if (accept) {
      c = (uint32_t)chrom[i];
      fl = flags[i];
    }
  }
  fc_jrec r;
  if (accept) r = fc::make_record(i, h.start, h.end, h.w2, h.w3, c, fl, e);
  fc::emit_block<256>(accept, r, e.n_recs, e.recs);
}

__global__ void key_hash_kernel(int64_t n, const fc_jrec* __restrict__ recs, uint64_t seed, uint64_t* __restrict__ h,
                                uint32_t* __restrict__ iota) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const fc_jrec& r = recs[i];
  h[i] = fc_key_hash(r.chrom, r.start, r.end, r.sk, seed);
  iota[i] = (uint32_t)i;
}

__global__ void idx_key_kernel(int64_t n, const fc_jrec* __restrict__ recs, uint64_t* __restrict__ k, uint32_t* __restrict__ iota) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  k[i] = recs[i].idx;
  iota[i] = (uint32_t)i;
}

__global__ void gather_kernel(int64_t n, const fc_jrec* __restrict__ recs, const uint32_t* __restrict__ perm,
                              fc_jrec* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* src = reinterpret_cast<const uint4*>(recs + perm[i]);
  uint4* dst = reinterpret_cast<uint4*>(out + i);
  dst[0] = src[0];
  dst[1] = src[1];
  dst[2] = src[2];
}

__device__ inline bool same_key(const fc_jrec& a, const fc_jrec& b) {
  return a.chrom == b.chrom && a.start == b.start && a.end == b.end && ((a.sk ^ b.sk) & 3u) == 0;
}

// head[i] = 1 when sorted record i starts a new junction; counts hash collisions (equal hash, different key)
__global__ void heads_kernel(int64_t n, const fc_jrec* __restrict__ s, const uint64_t* __restrict__ h_sorted,
                             uint64_t hmask, uint32_t* __restrict__ head, unsigned long long* __restrict__ collisions) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t hd = 1;
  if (i > 0) {
    bool same_h = ((h_sorted[i] ^ h_sorted[i - 1]) & hmask) == 0;  // only the sorted bits order the records
    bool same_k = same_key(s[i], s[i - 1]);
    if (same_h && !same_k) atomicAdd(collisions, 1ull);
    hd = same_k ? 0u : 1u;
  }
  head[i] = hd;
}

__global__ void acc_init_kernel(int64_t nj, JAcc* __restrict__ acc) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  JAcc a;
  a.n_spanned = 0;
  for (int k = 0; k < 4; ++k) a.cw[k] = a.cb[k] = 0;
  a.cw_other = a.cb_other = 0;
  a.max_ql = a.max_qr = -2147483647 - 1;
  a.min_dist = a.min_ov = a.min_nh = 0xFFFFFFFFu;
  a.n_frags = a.n_uniq = a.n_pal = 0;
  a.first_idx = ~0ull;
  a.seg_start = 0;
  a.pad = 0;
  acc[j] = a;
}

// one thread per sorted record; lanes of a warp that share a junction are combined before the atomics
__global__ void reduce_kernel(int64_t n, const fc_jrec* __restrict__ s, const uint32_t* __restrict__ seg_incl,
                              const uint32_t* __restrict__ head, JAcc* __restrict__ acc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool active = i < n;
  uint32_t seg = active ? seg_incl[i] - 1u : 0xFFFFFFFFu;
  unsigned amask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  fc_jrec r = s[i];
  if (head[i]) acc[seg].seg_start = (unsigned int)i;
  unsigned peers = __match_any_sync(amask, seg);
  int lane = threadIdx.x & 31;
  int leader = __ffs((int)peers) - 1;
  uint32_t den = (r.sk >> 8) & 0xFFu;
  int cls = den == 1 ? 0 : den == 2 ? 1 : den == 4 ? 2 : den == 8 ? 3 : 4;
  bool bridge = r.q_left != 0 && r.q_right != 0;
  // peers are contiguous lanes (records are sorted by junction): reduce with shuffles over the peer group
  unsigned cnt = __popc(peers);
  unsigned cw[5], cb[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    cw[k] = __popc(__ballot_sync(amask, cls == k) & peers);
    cb[k] = __popc(__ballot_sync(amask, cls == k && bridge) & peers);
  }
  int ql = r.q_left, qr = r.q_right;
  unsigned md = r.dist, mo = r.ov, mh = r.n_hits;
  unsigned long long fi = r.idx;
  // segmented butterfly inside the peer group
  int hi_lane = 31 - __clz((int)peers);
  for (int o = 1; o < 32; o <<= 1) {
    int src = lane + o;
    int ql2 = __shfl_down_sync(amask, ql, o);
    int qr2 = __shfl_down_sync(amask, qr, o);
    unsigned md2 = __shfl_down_sync(amask, md, o);
    unsigned mo2 = __shfl_down_sync(amask, mo, o);
    unsigned mh2 = __shfl_down_sync(amask, mh, o);
    unsigned long long fi2 = __shfl_down_sync(amask, fi, o);
    if (src <= hi_lane && ((peers >> src) & 1u)) {
      ql = max(ql, ql2);
      qr = max(qr, qr2);
      md = min(md, md2);
      mo = min(mo, mo2);
      mh = min(mh, mh2);
      fi = fi < fi2 ? fi : fi2;
    }
  }
  if (lane == leader) {
    JAcc* a = acc + seg;
    atomicAdd(&a->n_spanned, cnt);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (cw[k]) atomicAdd(&a->cw[k], cw[k]);
      if (cb[k]) atomicAdd(&a->cb[k], cb[k]);
    }
    if (cw[4]) atomicAdd(&a->cw_other, cw[4]);
    if (cb[4]) atomicAdd(&a->cb_other, cb[4]);
    atomicMax(&a->max_ql, ql);
    atomicMax(&a->max_qr, qr);
    atomicMin(&a->min_dist, md);
    atomicMin(&a->min_ov, mo);
    atomicMin(&a->min_nh, mh);
    atomicMin(&a->first_idx, fi);
  }
}

// ---- distinct counts with exact hash sets (128-bit entries, 128-bit CAS) ----------------------------------------
// insert (value, seg) into an open-addressing set; returns true when it was not present.  The whole element is the
// 128-bit entry, so membership is exact (no fingerprint collisions).
__device__ __forceinline__ bool set_insert(U128* table, unsigned long long mask, unsigned long long value, uint32_t seg) {
  const U128 mine{value, (unsigned long long)seg + 1ull};
  unsigned long long slot = fc_mix64(value ^ ((unsigned long long)seg * 0x9E3779B97F4A7C15ULL)) & mask;
  for (;;) {
    const ulonglong2 cur = __ldcg(reinterpret_cast<const ulonglong2*>(table + slot));
    if (cur.x == mine.lo && cur.y == mine.hi) return false;
    if (cur.x == 0ull && cur.y == 0ull) {
      const U128 old = cas128(table + slot, U128{0ull, 0ull}, mine);
      if (old.lo == 0ull && old.hi == 0ull) return true;
      if (old.lo == mine.lo && old.hi == mine.hi) return false;
    }
    slot = (slot + 1ull) & mask;
  }
}

// one thread per sorted record: is this the first time its read sequence / its fragment name is seen in its junction?
// Records are sorted by junction, so the lanes of a warp mostly share one junction: the per-junction counters are
// bumped once per (warp, junction) group.
__global__ void distinct_hash_kernel(int64_t n, const fc_jrec* __restrict__ s, const uint32_t* __restrict__ seg_incl,
                                     U128* __restrict__ tab_reads, U128* __restrict__ tab_names, unsigned long long mask,
                                     JAcc* __restrict__ acc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const unsigned amask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const uint32_t seg = seg_incl[i] - 1u;
  const unsigned long long rh = s[i].read_hash, qh = s[i].qname_hash;
  const bool new_read = set_insert(tab_reads, mask, rh, seg);
  const bool new_name = set_insert(tab_names, mask, qh, seg);
  const unsigned peers = __match_any_sync(amask, seg);
  const unsigned b_read = __ballot_sync(amask, new_read) & peers;
  const unsigned b_pal = __ballot_sync(amask, new_read && (rh & 1ull)) & peers;
  const unsigned b_name = __ballot_sync(amask, new_name) & peers;
  if ((int)(threadIdx.x & 31) == __ffs((int)peers) - 1) {
    JAcc* a = acc + seg;
    if (b_read) atomicAdd(&a->n_uniq, (unsigned)__popc(b_read));
    if (b_pal) atomicAdd(&a->n_pal, (unsigned)__popc(b_pal));
    if (b_name) atomicAdd(&a->n_frags, (unsigned)__popc(b_name));
  }
}

// Where the records of a call live: one run behind a counter (local emit), or -- multi-GPU -- `n_slices` slices of
// slice_cap records, slice s filled by source rank s with counts[s] records (emit_core.cuh: P2PView)
struct RecSrc {
  const fc_jrec* base;
  const unsigned long long* counts;  // n_slices words
  unsigned long long slice_cap;      // (n_slices == 1: the upper bound of the record count)
  int n_slices;
  unsigned long long* total_out;     // n_slices > 1: block 0 writes the exact record count here
};

// ======================================================================================================================
// Sort-free aggregation (default): ONE pass over the records.  Every junction lives in one 64-byte slot of an
// open-addressing table: the first sector holds its identity AND its extrema (one 32-byte load tells a record whether
// the slot is its junction and whether it improves any extreme -- it rarely does after a junction's first records), the
// second sector holds the counters.  A record of a junction that is not new costs: that one load, ONE reduction
// (n_spanned and the weight travel in one 64-bit word; everything that is unusual -- a read seen before, a fragment seen
// before, an anchor without uniqueness, a palindromic read -- is counted by its own, rare, reduction and subtracted at
// the end) and the insert of (read hash, junction) into one exact hash set for the distinct counts.  Weights are k/8
// (denominators 1,2,4,8), so fixed-point sums are exact in any order; inputs with another denominator fall back to the
// sort-based path below.
//
// The distinct set is cleared per call; the finish kernel zeroes exactly the slots it consumed, so the junction table is
// never cleared as a whole.
// ======================================================================================================================
struct alignas(64) JSlot {
  // sector 0: everything an ordinary record looks at or changes (maxima: all-zero = nothing yet = identity of max)
  unsigned long long klo;  // start | end << 32
  unsigned long long kf;   // KEY_OCC | chrom (20 bits) << 43 | (strand, kind) << 41 | first: FIRST_MAX - (smallest idx - idx base)
  unsigned long long c0;   // n_spanned | 8 * n_weighted << 32
  unsigned long long ext;  // (q_left + 32768) | (q_right + 32768) << 16 | (255 - ov) << 32 | (255 - dist) << 40 | (65535 - n_hits) << 48
  // sector 1: what few records touch
  unsigned long long c1;   // 8 * weight of records that are no unique bridge | records whose fragment name is not new << 32
  unsigned long long c2;   // 2 per record whose read was seen before + 1 per new palindromic read
  unsigned int sig;        // 12-bit signal code (same for every record of the junction)
  unsigned int pad;
  unsigned long long spare;
};
static_assert(sizeof(JSlot) == 64 && offsetof(JSlot, c0) == 16 && offsetof(JSlot, c1) == 32, "JSlot layout");

constexpr unsigned long long KEY_OCC = 1ull << 63;
constexpr int FIRST_BITS = 41;  // record positions relative to the call's idx base: 2.2e12 (a native ingest numbers 64 per fragment)
constexpr int CHROM_BITS = 20;  // chromosome numbers (a call with more falls back to the sort-based path)
constexpr unsigned long long FIRST_MAX = (1ull << FIRST_BITS) - 1ull;
constexpr long long FUSED_MAX_RECORDS = 1ll << 28;  // keeps 8 * n_spanned and the slot numbers inside their fields
constexpr uint32_t SK_NAME_KNOWN = FC_SK_NAME_KNOWN, SK_NAME_DUP = FC_SK_NAME_DUP;  // the emitter already knows whether the fragment is new to the junction

// counters (32-bit words at counters + 8): [0] junctions listed, [1] records with another denominator,
// [2] list overflow / records outside a declared idx range, [3] junctions, [4] set entries that found their partition
// full, [5] partitions whose shared-memory set ran full (either: the call is repeated with the global set),
enum { FC_N_ALLOC = 0, FC_N_OTHER = 1, FC_N_OVERFLOW = 2, FC_N_JUNC = 3, FC_N_PART_OVF = 4, FC_N_SET_FULL = 5, FC_N_TABLE_FULL = 6 };
constexpr int FC_N_CTR = 8;
// [6] records that found no place for their junction within SLOT_MAX_PROBES slots of a junction table that was sized after
// the last call's junction count: the call is repeated with the table sized after the record count
constexpr int SLOT_MAX_PROBES = 64;

__device__ __forceinline__ unsigned long long ext_pack(int q_left, int q_right, unsigned dist, unsigned ov, unsigned n_hits) {
  const unsigned lo = (unsigned)(q_left + 32768) | ((unsigned)(q_right + 32768) << 16);
  const unsigned hi = (255u - ov) | ((255u - dist) << 8) | ((65535u - n_hits) << 16);
  return (unsigned long long)lo | ((unsigned long long)hi << 32);
}
// field-wise maximum of two packed extrema words
__device__ __forceinline__ unsigned long long ext_max(unsigned long long a, unsigned long long b) {
  const unsigned alo = (unsigned)a, blo = (unsigned)b, ahi = (unsigned)(a >> 32), bhi = (unsigned)(b >> 32);
  const unsigned lo = __vmaxu2(alo, blo);
  const unsigned hi = (__vmaxu2(ahi, bhi) & 0xFFFF0000u) | (__vmaxu4(ahi, bhi) & 0x0000FFFFu);
  return (unsigned long long)lo | ((unsigned long long)hi << 32);
}
// does x beat cur in any field?  (the usual answer is no; ext_max is only worth computing after a yes)
__device__ __forceinline__ bool ext_improves(unsigned long long cur, unsigned long long x) {
  const unsigned clo = (unsigned)cur, xlo = (unsigned)x, chi = (unsigned)(cur >> 32), xhi = (unsigned)(x >> 32);
  return (xlo & 0xFFFFu) > (clo & 0xFFFFu) || (xlo >> 16) > (clo >> 16) || (xhi & 0xFFu) > (chi & 0xFFu) ||
         ((xhi >> 8) & 0xFFu) > ((chi >> 8) & 0xFFu) || (xhi >> 16) > (chi >> 16);
}
// kf / x: identity + first position and extrema of one record (or of one shared-memory entry); cur_kf / cur_x: what a
// (possibly stale: maxima only grow, so an old value can only cause a needless attempt) look at the slot showed
__device__ __forceinline__ void extrema_to_global(JSlot* s, unsigned long long kf, unsigned long long x, unsigned long long cur_kf,
                                                  unsigned long long cur_x) {
  if (kf > cur_kf) atomicMax(&s->kf, kf);  // (the identity bits are the same for every record of the junction)
  if (!ext_improves(cur_x, x)) return;
  unsigned long long want = ext_max(cur_x, x);
  while (want != cur_x) {
    const unsigned long long old = atomicCAS(&s->ext, cur_x, want);
    if (old == cur_x) break;
    cur_x = old;
    want = ext_max(cur_x, x);
  }
}

// Exact set of (value, tag) pairs, insert only.  Probe sequence of an element: first the slot given by the value alone
// (known before the junction is, so its sector can be prefetched while the junction table is read -- an atomic that
// misses in L2 is far slower than one that hits), then linear probing from a slot that also depends on the tag (a
// read sequence that supports very many junctions does not build one long cluster).
__device__ __forceinline__ unsigned long long set_first_slot(unsigned long long v, unsigned long long mask) {
  return fc_mix64(v) & mask;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// insert (v1, t1) and, when with2, (v2, t2); both probe sequences advance together so that their round trips overlap.
// The CAS goes first (no load): most inserts find an empty slot, and the returned old entry tells the rest.  Probe
// sequence of an element: the slot given by its value alone, that slot's neighbour in the same 32-byte sector (the
// sector is in L2 by then), then linear probing from a slot that also depends on the tag.
struct SetProbe {
  unsigned long long s;
  int k;
  __device__ __forceinline__ void next(unsigned long long v, unsigned long long t, unsigned long long mask) {
    if (k == 0)
      s ^= 1ull;
    else if (k == 1)
      s = fc_mix64(v ^ (t * 0x9E3779B97F4A7C15ULL)) & mask;
    else
      s = (s + 1ull) & mask;
    ++k;
  }
};
// (entries never change once written, so a look through L1 is as good as one at L2; an EMPTY seen there may be stale: the
// compare-and-swap then returns the real owner)
__device__ __forceinline__ U128 set_peek(const U128* p) {
  U128 v;
  asm volatile("ld.global.ca.v2.u64 {%0,%1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(p));
  return v;
}
__device__ __forceinline__ void set_insert2(U128* table, unsigned long long mask, unsigned long long v1, unsigned long long t1,
                                            unsigned long long s1, bool with2, unsigned long long v2, unsigned long long t2,
                                            unsigned long long s2, bool peek, bool& new1, bool& new2) {
  bool d1 = false, d2 = !with2;
  SetProbe p1{s1, 0}, p2{s2, 0};
  new1 = new2 = false;
  while (!(d1 && d2)) {
    // look first when the set is larger than L2: in a popular junction most reads have been seen before, an element that
    // is already there is found by a (cacheable) load, and a load that misses to DRAM is cheaper than an atomic that does
    // (measured, 50 M records: 3.59 -> 3.24 ms; 54 M records over 10 M junctions: 12.2 -> 8.5 ms).  A set that fits L2 has
    // had its slots prefetched and goes straight to the atomic (the look costs 15 % there: one more round trip).
    U128 o1{0ull, 0ull}, o2{0ull, 0ull};
    if (peek && !d1) o1 = set_peek(table + p1.s);
    if (peek && !d2) o2 = set_peek(table + p2.s);
    if (!d1 && o1.lo == 0ull && o1.hi == 0ull) o1 = cas128(table + p1.s, U128{0ull, 0ull}, U128{v1, t1});
    else if (!d1 && !(o1.lo == v1 && o1.hi == t1)) o1.lo = ~v1;  // (somebody else's: probe on; cannot equal mine or be empty)
    if (!d2 && o2.lo == 0ull && o2.hi == 0ull) o2 = cas128(table + p2.s, U128{0ull, 0ull}, U128{v2, t2});
    else if (!d2 && !(o2.lo == v2 && o2.hi == t2)) o2.lo = ~v2;
    if (!d1) {
      if (o1.lo == 0ull && o1.hi == 0ull)
        new1 = d1 = true;
      else if (o1.lo == v1 && o1.hi == t1)
        d1 = true;
      else
        p1.next(v1, t1, mask);
    }
    if (!d2) {
      if (o2.lo == 0ull && o2.hi == 0ull)
        new2 = d2 = true;
      else if (o2.lo == v2 && o2.hi == t2)
        d2 = true;
      else
        p2.next(v2, t2, mask);
    }
  }
}

// ---- Partitioned distinct counts (inputs whose set would not fit L2) -------------------------------------------------
// A global hash set costs every record one random DRAM sector (and the part sustains only ~30 G of those per second).
// Instead, pass 1 turns every (read, junction) -- and every (fragment name, junction) that the scan kernel could not
// settle -- into a 16-byte entry {64-bit key, junction | flags} and appends it to one of n_parts partitions chosen by
// an independent hash of the same pair; pass 2 (distinct_parts_kernel) walks each partition with an exact set in SHARED
// memory and adds what it finds to the junctions' counters.  Equal pairs meet in one partition, so the counts are the
// ones of the global set; the distinct elements spread evenly whatever the popularity of the junctions is, and the
// traffic is one sequential 16-byte write and read per record.  Two pairs are taken for equal when their 64-bit keys
// AND partitions agree: ~2^-64 per pair of elements of one partition (DESIGN.md section 2 has the bound per run).
// Copies of an element (the same read sequence on a popular junction comes thousands of times) would crowd its
// partition: each CTA of pass 1 keeps the keys it sent last in a small direct-mapped cache and counts a hit as the
// duplicate it is.
struct PartView {
  uint4* ent;           // n_parts x pcap entries, the partitions interleaved in chunks of PART_CHUNK entries (part_slot)
  unsigned int* cur;    // fill count of partition p at cur[8 * p] (own sector each); all zero between calls
  unsigned int n_parts;
  unsigned int pcap;    // a multiple of PART_CHUNK
};
// Entry `pos` of partition `part` lives in chunk pos / PART_CHUNK of that partition, and chunk c of all partitions lies side by
// side.  The partitions fill at the same pace (the elements spread evenly), so the appends of any moment fall into one row
// of chunks -- n_parts x 4 KB, within the reach of the TLB -- instead of all over the gigabytes of the partition space:
// measured with profiles/tools/scatter_rate.cu, an append (returning atomic + dependent 16-byte store) costs an SM 8.5 cycles
// with every partition in a region of its own, 4.5 interleaved.
constexpr unsigned PART_CHUNK_BITS = 8, PART_CHUNK = 1u << PART_CHUNK_BITS;
__device__ __forceinline__ size_t part_slot(const PartView& pv, unsigned int part, unsigned int pos) {
  return (((size_t)(pos >> PART_CHUNK_BITS) * pv.n_parts + part) << PART_CHUNK_BITS) + (pos & (PART_CHUNK - 1u));
}
constexpr unsigned PART_NAME = 1u << 31, PART_PALIN = 1u << 30, PART_JID = (1u << 30) - 1u;  // (slot numbers stay below 2^30: FUSED_MAX_RECORDS)
constexpr int RECENT_SETS = 2048;     // per-CTA cache of keys sent before: 2 ways x 8 bytes per set (32 KB, dynamic shared memory)
constexpr int PSET_ENTRIES = 8192;    // shared-memory set of pass 2 (64 KB: three CTAs per SM)
constexpr int PSET_TARGET = 2600;     // elements per partition the host aims at (load 0.32; 0.64 when every name goes through it too)
constexpr int PART_THREADS = 512;

// 16 bytes of shared memory in one instruction, not cached in registers across calls
__device__ __forceinline__ ulonglong2 lds128(const void* p) {
  ulonglong2 v;
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
  return v;
}

// key (never 0) and partition of the pair (value, tag): two independent 64-bit mixes of the pair.  (The values are hashes
// already, but the caller's: a cheaper key such as v ^ t * odd makes (v, t) and (t, v) one element when somebody's hashes
// are small multiples of that constant -- and the kernel's time does not depend on its instruction count, section 4.2.)
__device__ __forceinline__ unsigned long long part_key(unsigned long long v, unsigned long long t, unsigned int n_parts, unsigned int& part) {
  const unsigned long long tm = fc_mix64(t * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL);
  const unsigned long long k = fc_mix64(v ^ tm);
  const unsigned long long m2 = fc_mix64((v + 0xD1B54A32D192ED03ULL) * 0xA24BAED4963EE407ULL ^ (tm >> 29) ^ (tm << 35));
  part = __umulhi((unsigned int)(m2 >> 32), n_parts);
  return k ? k : 1ull;
}
// claim a place in the key's partition unless this CTA has sent the same key before and still remembers it (returns
// false: a repeat).  The memory is a two-way cache with a protected way: a key enters way 1 and moves to way 0 when it comes
// again, so the few thousand keys that make up most of the repeats are not pushed out by the stream of keys that come
// once.  Every value in the cache is a key that was sent, whatever the races between the lanes do to it.
__device__ __forceinline__ bool part_claim(const PartView& pv, unsigned long long* recent, unsigned long long k, unsigned int part, unsigned int& pos) {
  volatile unsigned long long* rc = recent + 2u * ((unsigned int)(k >> 20) & (RECENT_SETS - 1));
  const ulonglong2 ways = lds128(recent + 2u * ((unsigned int)(k >> 20) & (RECENT_SETS - 1)));  // (both ways in one 16-byte load)
  const unsigned long long w0 = ways.x, w1 = ways.y;
  if (w0 == k) return false;
  if (w1 == k) {
    rc[0] = k;
    rc[1] = w0;
    return false;
  }
  rc[1] = k;
  pos = atomicAdd(pv.cur + 8u * part, 1u);  // (nobody looks at the answer before part_store: the round trip stays off the critical path)
  return true;
}
__device__ __forceinline__ void part_store(const PartView& pv, unsigned long long k, unsigned int part, unsigned int pos, unsigned int word,
                                           unsigned int* ctr) {
  if (pos < pv.pcap)
    pv.ent[part_slot(pv, part, pos)] = make_uint4((unsigned int)k, (unsigned int)(k >> 32), word, 0u);
  else
    atomicAdd(&ctr[FC_N_PART_OVF], 1u);
}

// A junction that collects a few per cent of all reads (expression is heavy-tailed) would otherwise put all of its
// updates on one L2 sector, and an L2 slice retires about one request per clock for one sector (measured: with a
// Zipf(1) popularity the direct version spends 4x the time of the uniform case).  So the lanes of a warp that share a
// junction are combined first, and a junction that two lanes of one warp have shared once enters a shared-memory table
// of the CTA, is accumulated there (found with a plain load) and flushed at the end of the CTA's chunk of records;
// the bulk of the distinct junctions never meet that condition, go to global memory directly and never touch the table.
// A CTA walks its chunk tile by tile WITHOUT barriers in between (whichever way a record goes, it ends up in the
// junction's slot), so the warps hide each other's memory round trips.  Measured on the B200 (config 3, 44 M records):
// without the table 9.96 ms, with it 3.6 ms; a count-min sketch as a second admission rule changed nothing.
constexpr int ACC_THREADS = 512;
// History of what was measured on the B200 (configs 3 and 5) while the kernel still used one global set for every input:
// 3 CTAs per SM (40 registers, spills), chunks of 16 instead of 30 tiles, 8 instead of 2 probes of the shared-memory table
// and an L2 prefetch of the next tile's records all changed the kernel time by < 3 %; two L2-resident bit filters in front
// of the read set (only pairs whose bit is hit twice reach the exact set in a second pass) gained nothing because reads
// repeat so often in popular junctions.  With the set partitioned (above) the kernel is a chain of three dependent
// round trips per record -- record, junction slot, partition cursor -- and what pays is taking them off the critical
// path costs more (registers, shared-memory traffic) than it hides: a three-deep software pipeline over the tiles of a chunk
// (record, then slot sector one tile ahead) measured 6 % slower than the plain loop below.
#define FC_ACC_MIN_CTAS 2
#define FC_HOT_PROBES 2
constexpr int ACC_MAX_TILES = 30;  // tiles of ACC_THREADS records per chunk (between two flushes of the shared-memory table)
constexpr int HOT_ENTRIES = 512;
constexpr int HOT_BITS = 9;
struct alignas(32) HotEntry {  // one junction of the CTA's table: everything a record looks at or changes in two 16-byte halves
  unsigned int tag;                // junction slot + 1, 0 = free
  unsigned int c0;                 // n_spanned | 8 * weight << 14 (a chunk holds fewer than 16384 records of weight <= 1)
  unsigned int c1;                 // names seen before | 8 * non-bridge weight << 14
  unsigned int c2;
  unsigned long long ext;          // packed extrema, as JSlot.ext
  unsigned long long first_inv;    // identity | first position, as JSlot.kf
};
struct HotTable {
  HotEntry e[HOT_ENTRIES];
};
static_assert(ACC_MAX_TILES * ACC_THREADS < (1 << 14) && ACC_MAX_TILES * ACC_THREADS * 8 < (1 << 18), "hot-table fields");

// entry of junction `jid` in the CTA's table; `insert`: claim a free entry when it has none (-1: none / table crowded).
// A plain load finds a junction that is already there -- the usual case for a popular one -- without an atomic.
__device__ __forceinline__ int hot_find_or_insert(HotTable& t, unsigned int jid, bool insert) {
  unsigned int h = (jid * 2654435761u) >> (32 - HOT_BITS);
#pragma unroll 1
  for (int probe = 0; probe < FC_HOT_PROBES; ++probe) {  // (2 probes: a miss is the common case and must stay cheap)
    unsigned int cur = *reinterpret_cast<volatile unsigned int*>(&t.e[h].tag);
    if (cur == 0u) {
      if (!insert) return -1;
      cur = atomicCAS(&t.e[h].tag, 0u, jid + 1u);
      if (cur == 0u) return (int)h;
    }
    if (cur == jid + 1u) return (int)h;
    h = (h + 1u) & (HOT_ENTRIES - 1);
  }
  return -1;  // crowded: the group goes to global memory directly
}

// one 32-byte sector (identity + extrema of a slot)
__device__ __forceinline__ void ld_sector(const JSlot* s, unsigned long long (&v)[4]) {
  asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(s));
}

// where record i of the call lives (several slices: the concatenation of what the source ranks sent)
__device__ __forceinline__ const uint4* rec_ptr(const RecSrc& src, int64_t i) {
  int64_t rec_at = i;
  if (src.n_slices > 1) {
    int64_t before = 0;
#pragma unroll 1
    for (int sl = 0; sl < src.n_slices; ++sl) {
      const int64_t c = (int64_t)min(src.counts[sl], src.slice_cap);
      if (i >= before && i < before + c) rec_at = (int64_t)sl * (int64_t)src.slice_cap + (i - before);
      before += c;
    }
  }
  return reinterpret_cast<const uint4*>(src.base + rec_at);
}
// first word of a record (chrom, start, end, sk) -> the slot's key words and where its probe sequence starts
__device__ __forceinline__ void key_of(const uint4& r0, unsigned long long& klo, unsigned long long& kid) {
  klo = (unsigned long long)r0.y | ((unsigned long long)r0.z << 32);
  kid = KEY_OCC | ((unsigned long long)(r0.x & ((1u << CHROM_BITS) - 1u)) << (FIRST_BITS + 2)) | ((unsigned long long)(r0.w & 3u) << FIRST_BITS);
}
__device__ __forceinline__ unsigned long long slot_of(unsigned long long klo, unsigned long long kid, unsigned long long kmask) {
  return fc_mix64(klo ^ ((kid >> FIRST_BITS) * 0x9E3779B97F4A7C15ULL)) & kmask;
}

// SETS: 0 = one global set (inputs whose set fits L2), 1 = partitioned (PartView; distinct_parts_kernel follows).
template <int SETS>
__global__ void __launch_bounds__(ACC_THREADS, FC_ACC_MIN_CTAS) fused_accumulate_kernel(RecSrc src, int chunk_tiles, int prefetch, int max_probes, unsigned long long idx_base, JSlot* __restrict__ slots,
                                                                  unsigned long long kmask, U128* __restrict__ sets,
                                                                  unsigned long long smask, PartView pv, unsigned int* __restrict__ list,
                                                                  unsigned int lcap, unsigned int* __restrict__ ctr,
                                                                  uint4* __restrict__ flag4, int64_t n_flag4,
                                                                  uint32_t* __restrict__ tile_count, int64_t n_tiles) {
  __shared__ HotTable hot;
  extern __shared__ unsigned long long recent[];  // SETS == 1: 2 * RECENT_SETS words
  if (SETS == 1) {
    for (int e = threadIdx.x; e < 2 * RECENT_SETS; e += ACC_THREADS) recent[e] = 0ull;  // (the first barrier below orders this)
  }
  // the rank flags and tile counters of the finish pass are cleared on the way
  {
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gs = (int64_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t k = gi; k < n_flag4; k += gs) flag4[k] = z;
    for (int64_t k = gi; k < n_tiles; k += gs) tile_count[k] = 0u;
  }
  // records of the call: slice s holds records [prefix(s), prefix(s+1)) (a counter may have counted past its capacity)
  int64_t n = 0;
#pragma unroll 1
  for (int sl = 0; sl < src.n_slices; ++sl) n += (int64_t)min(src.counts[sl], src.slice_cap);
  if (src.total_out && blockIdx.x == 0 && threadIdx.x == 0) *src.total_out = (unsigned long long)n;
  const int64_t chunk_recs = (int64_t)chunk_tiles * ACC_THREADS;
  const unsigned lane = threadIdx.x & 31;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

#pragma unroll 1
  for (int64_t c0 = (int64_t)blockIdx.x * chunk_recs; c0 < n; c0 += (int64_t)gridDim.x * chunk_recs) {
    for (int e = threadIdx.x; e < HOT_ENTRIES; e += ACC_THREADS) {
      uint4* z = reinterpret_cast<uint4*>(&hot.e[e]);
      z[0] = z[1] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    // (a junction table that turned out too small: the call is going to be repeated, the chunks left have nothing to do)
    const bool give_up = *reinterpret_cast<volatile unsigned int*>(&ctr[FC_N_TABLE_FULL]) != 0u;
    int64_t i = c0 + threadIdx.x;
#pragma unroll 1
    for (int t = 0; t < chunk_tiles; ++t, i += ACC_THREADS) {
      const bool active = i < n && !give_up;
      const unsigned amask = __ballot_sync(0xffffffffu, active);
      if (!active) continue;  // (trailing lanes of the last tile; the masks below name the active lanes only)
      const uint4* rp = rec_ptr(src, i);
      const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
      // fc_jrec: chrom,start,end,sk | idx, read_hash | qname_hash, q_left,q_right, n_hits,dist,ov
      const unsigned sk = r0.w;
      const unsigned long long idx = (unsigned long long)r1.x | ((unsigned long long)r1.y << 32);
      const unsigned long long read_hash = (unsigned long long)r1.z | ((unsigned long long)r1.w << 32);
      const unsigned long long qname_hash = (unsigned long long)r2.x | ((unsigned long long)r2.y << 32);
      const bool name_known = (sk & SK_NAME_KNOWN) != 0u;
      unsigned long long klo = 0ull, kid = 0ull, kf = 0ull, cur_kf = 0ull, cur_x = 0ull, s_read = 0ull, s_name = 0ull;
      bool fresh = false, sent_r = false, sent_n = false;
      unsigned int part_r = 0u, part_n = 0u, pos_r = 0u, pos_n = 0u;
      unsigned long long key_r = 0ull, key_n = 0ull;
      // names and reads share the set: the name's value is salted so that equal hashes of the two kinds stay apart
      if (SETS == 0) {
        s_read = set_first_slot(read_hash, smask);
        if (prefetch) prefetch_l2(sets + s_read);
        if (!name_known) {
          s_name = set_first_slot(~qname_hash, smask);
          if (prefetch) prefetch_l2(sets + s_name);
        }
      }

      // ---- the junction's slot
      key_of(r0, klo, kid);
      const unsigned long long rel = idx - idx_base;
      // (a position or chromosome number beyond the slot's fields: the call falls back to the sort-based path)
      if ((rel >> FIRST_BITS) != 0ull || (r0.x >> CHROM_BITS) != 0u) atomicAdd(&ctr[FC_N_OTHER], 1u);
      kf = kid | (FIRST_MAX - (rel & FIRST_MAX));
      unsigned long long slot = slot_of(klo, kid, kmask);
      bool dead = false;  // no place found: the lane goes along (the warp-wide steps below need it) without touching any state
      for (int probes = 0;; ++probes) {
        if (probes == max_probes) {  // (only a table sized after the last batch has a limit: the full table always has room)
          atomicAdd(&ctr[FC_N_TABLE_FULL], 1u);
          dead = true;
          break;
        }
        // L1-cached look: the identity never changes once written and the maxima only grow, so a cached copy is as
        // good as the one in L2 (the slot of a popular junction is read by thousands of threads); a cached EMPTY may be
        // stale: the compare-and-swap below then returns the real owner
        unsigned long long v[4];
        ld_sector(slots + slot, v);
        if (v[0] == 0ull && v[1] == 0ull) {
          const U128 old = cas128(reinterpret_cast<U128*>(slots + slot), U128{0ull, 0ull}, U128{klo, kf});
          if (old.lo == 0ull && old.hi == 0ull) {
            fresh = true;  // a new junction (listed for the finish pass below)
            cur_kf = kf;
            break;
          }
          v[0] = old.lo;
          v[1] = old.hi;
          v[3] = 0ull;  // extrema unknown: "nothing yet" is a valid (stale) view
        }
        if (v[0] == klo && (v[1] >> FIRST_BITS) == (kid >> FIRST_BITS)) {
          cur_kf = v[1];
          cur_x = v[3];
          break;
        }
        slot = (slot + 1ull) & kmask;
      }
      // ---- partitioned distinct counts: claim the places in the partitions now, the answers are used at the very end
      if (SETS == 1 && !dead) {
        const unsigned long long tag = (unsigned long long)(unsigned int)slot + 1ull;
        key_r = part_key(read_hash, tag, pv.n_parts, part_r);
        sent_r = part_claim(pv, recent, key_r, part_r, pos_r);
        if (!name_known) {
          key_n = part_key(~qname_hash, tag | (1ull << 32), pv.n_parts, part_n);
          sent_n = part_claim(pv, recent, key_n, part_n, pos_n);
        }
      }
      // slot number = the junction's id in this call (a lane without a slot: a number no slot has, its own in the warp)
      const unsigned int jid = dead ? 0xFFFFFF00u | lane : (unsigned int)slot;
      JSlot* a = slots + (dead ? 0u : jid);
      const unsigned long long tag = (unsigned long long)jid + 1ull;  // never 0: no set entry is all-zero
      {
        // new junctions of the warp: one atomic for all of them claims their places in the list
        const unsigned fm = __ballot_sync(amask, fresh);
        if (fm) {
          const int fl = __ffs((int)fm) - 1;
          unsigned int base = 0;
          if ((int)lane == fl) base = atomicAdd(&ctr[FC_N_ALLOC], (unsigned int)__popc(fm));
          base = __shfl_sync(amask, base, fl);
          if (fresh) {
            const unsigned int pos = base + (unsigned int)__popc(fm & ((1u << lane) - 1u));
            if (pos < lcap)
              list[pos] = (unsigned int)slot;
            else
              atomicAdd(&ctr[FC_N_OVERFLOW], 1u);
            slots[slot].sig = (sk >> 16) & 0xFFFu;
          }
        }
      }

      // ---- lanes of the warp that share the junction are combined; is it in the CTA's table (or does it belong there)?
      const unsigned peers = __match_any_sync(amask, jid);
      const unsigned group = __popc(peers);
      const int leader = __ffs((int)peers) - 1;
      const bool lead = (int)lane == leader;
      int he = -1;
      if (lead) he = hot_find_or_insert(hot, jid, group >= 2u);
      he = __shfl_sync(peers, he, leader);

      // ---- extrema
      const int q_left = (int)(short)(r2.z & 0xFFFFu), q_right = (int)(short)(r2.z >> 16);
      const unsigned n_hits = r2.w & 0xFFFFu, dist = (r2.w >> 16) & 0xFFu, ov = r2.w >> 24;
      const unsigned long long x = ext_pack(q_left, q_right, dist, ov, n_hits);
      if (he >= 0) {
        // (one 16-byte load shows both words; they only grow, so a stale view costs a needless attempt at worst)
        const ulonglong2 seen = lds128(&hot.e[he].ext);
        if (kf > seen.y) atomicMax(&hot.e[he].first_inv, kf);
        unsigned long long cx = seen.x, want = cx;
        if (ext_improves(cx, x)) want = ext_max(cx, x);
        while (want != cx) {  // (rare once the entry has seen a few records)
          const unsigned long long old = atomicCAS(&hot.e[he].ext, cx, want);
          if (old == cx) break;
          cx = old;
          want = ext_max(cx, x);
        }
      } else if (!dead) {
        extrema_to_global(a, kf, x, cur_kf, cur_x);
      }

      // ---- distinct reads / fragment names of the junction
      bool new_read = false, new_name = false;
      if (SETS == 0) {
        if (!dead) set_insert2(sets, smask, read_hash, tag, s_read, !name_known, qname_hash, tag | (1ull << 32), s_name, !prefetch, new_read, new_name);
      } else {
        // pass 2 decides; a repeat of a key this CTA has sent before is counted as the duplicate it is right here
        new_read = sent_r;
        new_name = sent_n;
      }
      const bool dup_name = name_known ? (sk & SK_NAME_DUP) != 0u : !new_name;

      // ---- counters
      const unsigned den = (sk >> 8) & 0xFFu;
      const int cls = den == 1 ? 0 : den == 2 ? 1 : den == 4 ? 2 : den == 8 ? 3 : 4;
      if (cls == 4) atomicAdd(&ctr[FC_N_OTHER], 1u);
      const unsigned fx = cls < 4 ? (8u >> cls) : 0u;
      const bool bridge = q_left != 0 && q_right != 0;
      const unsigned nb = bridge ? 0u : fx;
      // (bit 0 of the hash flags a palindromic read; partitioned: pass 2 adds what a read that was sent contributes)
      const unsigned c2 = new_read ? (SETS == 0 ? (unsigned)(read_hash & 1ull) : 0u) : 2u;
      if (__all_sync(amask, group == 1u)) {
        // no two lanes of the warp share a junction (the usual case)
        if (he >= 0) {
          atomicAdd(&hot.e[he].c0, 1u | (fx << 14));
          if (nb | (unsigned)dup_name) atomicAdd(&hot.e[he].c1, (unsigned)dup_name | (nb << 14));
          if (c2) atomicAdd(&hot.e[he].c2, c2);
        } else if (!dead) {
          atomicAdd(&a->c0, 1ull | ((unsigned long long)fx << 32));
          if (nb | (unsigned)dup_name) atomicAdd(&a->c1, (unsigned long long)nb | ((unsigned long long)dup_name << 32));
          if (c2) atomicAdd(&a->c2, (unsigned long long)c2);
        }
      } else {
        unsigned w = 0, b = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          w += (8u >> k) * __popc(__ballot_sync(amask, cls == k) & peers);
          b += (8u >> k) * __popc(__ballot_sync(amask, cls == k && !bridge) & peers);
        }
        const unsigned dups = __popc(__ballot_sync(amask, dup_name) & peers);
        const unsigned c2s = 2u * __popc(__ballot_sync(amask, c2 == 2u) & peers) + __popc(__ballot_sync(amask, c2 == 1u) & peers);
        if (lead) {
          if (he >= 0) {
            atomicAdd(&hot.e[he].c0, group | (w << 14));
            if (b | dups) atomicAdd(&hot.e[he].c1, dups | (b << 14));
            if (c2s) atomicAdd(&hot.e[he].c2, c2s);
          } else if (!dead) {
            atomicAdd(&a->c0, (unsigned long long)group | ((unsigned long long)w << 32));
            if (b | dups) atomicAdd(&a->c1, (unsigned long long)b | ((unsigned long long)dups << 32));
            if (c2s) atomicAdd(&a->c2, (unsigned long long)c2s);
          }
        }
      }
      if (SETS == 1) {
        if (sent_r) part_store(pv, key_r, part_r, pos_r, jid | ((read_hash & 1ull) ? PART_PALIN : 0u), ctr);
        if (sent_n) part_store(pv, key_n, part_n, pos_n, jid | PART_NAME, ctr);
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < HOT_ENTRIES; e += ACC_THREADS) {
      const HotEntry h = hot.e[e];
      if (h.tag == 0u) continue;
      JSlot* a = slots + (h.tag - 1u);
      extrema_to_global(a, h.first_inv, h.ext, __ldcg(&a->kf), __ldcg(&a->ext));
      const unsigned c0v = h.c0, c1v = h.c1, c2v = h.c2;
      if (c0v) atomicAdd(&a->c0, (unsigned long long)(c0v & 0x3FFFu) | ((unsigned long long)(c0v >> 14) << 32));
      if (c1v) atomicAdd(&a->c1, (unsigned long long)(c1v >> 14) | ((unsigned long long)(c1v & 0x3FFFu) << 32));
      if (c2v) atomicAdd(&a->c2, (unsigned long long)c2v);
    }
    __syncthreads();  // the table is re-initialised for the next chunk
  }
}

// Pass 2 of the partitioned distinct counts: one CTA per partition at a time, an exact set of the partition's 64-bit keys
// in shared memory.  What an entry adds to its junction (JSlot.c2: 2 per read seen before + 1 per new palindromic read;
// JSlot.c1 high word: fragment names seen before) is collected per junction in a small shared-memory table first -- the
// repeats belong to the popular junctions -- and flushed once per partition.  The partition's fill count is zeroed on
// the way out, so the buffers are clean for the next call.
constexpr int PADD_ENTRIES = 512;   // (64 KB of set + 6 KB of this table: three CTAs per SM)
__global__ void __launch_bounds__(PART_THREADS) distinct_parts_kernel(PartView pv, JSlot* __restrict__ slots, unsigned int* __restrict__ ctr) {
  extern __shared__ unsigned long long pset[];  // PSET_ENTRIES
  __shared__ unsigned int a_tag[PADD_ENTRIES], a_c2[PADD_ENTRIES], a_c1[PADD_ENTRIES];
#pragma unroll 1
  for (unsigned int p = blockIdx.x; p < pv.n_parts; p += gridDim.x) {
    const unsigned int fill = pv.cur[8u * p];
    if (fill == 0u) continue;  // (block-uniform)
    const unsigned int n = fill < pv.pcap ? fill : pv.pcap;
    {
      ulonglong2* z = reinterpret_cast<ulonglong2*>(pset);
      for (int e = threadIdx.x; e < PSET_ENTRIES / 2; e += PART_THREADS) z[e] = make_ulonglong2(0ull, 0ull);
      for (int e = threadIdx.x; e < PADD_ENTRIES; e += PART_THREADS) a_tag[e] = a_c2[e] = a_c1[e] = 0u;
    }
    __syncthreads();
#pragma unroll 1
    for (unsigned int i = threadIdx.x; i < n; i += PART_THREADS) {
      const uint4 e = __ldcs(pv.ent + part_slot(pv, p, i));
      const unsigned long long k = (unsigned long long)e.x | ((unsigned long long)e.y << 32);
      unsigned int h = (unsigned int)(k >> 40) & (PSET_ENTRIES - 1);
      bool fresh = false, placed = false;
#pragma unroll 1
      for (int probe = 0; probe < PSET_ENTRIES; ++probe) {
        const unsigned long long cur = atomicCAS(&pset[h], 0ull, k);  // (no look first: most keys are new and most slots free)
        if (cur == 0ull) {
          fresh = placed = true;
          break;
        }
        if (cur == k) {
          placed = true;
          break;
        }
        h = (h + 1u) & (PSET_ENTRIES - 1);
      }
      if (!placed) {  // the set is full: the host repeats the call with the global set
        atomicAdd(&ctr[FC_N_SET_FULL], 1u);
        continue;
      }
      const bool is_name = (e.z & PART_NAME) != 0u;
      const unsigned int add2 = is_name ? 0u : (fresh ? ((e.z & PART_PALIN) ? 1u : 0u) : 2u);
      const unsigned int add1 = (is_name && !fresh) ? 1u : 0u;
      if ((add2 | add1) == 0u) continue;
      const unsigned int jid = e.z & PART_JID;
      unsigned int s = (jid * 2654435761u) >> (32 - 9);
      bool done = false;
#pragma unroll 1
      for (int probe = 0; probe < 4 && !done; ++probe) {
        unsigned int cur = *reinterpret_cast<volatile unsigned int*>(&a_tag[s]);
        if (cur == 0u) cur = atomicCAS(&a_tag[s], 0u, jid + 1u);
        if (cur == 0u || cur == jid + 1u) {
          if (add2) atomicAdd(&a_c2[s], add2);
          if (add1) atomicAdd(&a_c1[s], add1);
          done = true;
        }
        s = (s + 1u) & (PADD_ENTRIES - 1);
      }
      if (!done) {
        JSlot* a = slots + jid;
        if (add2) atomicAdd(&a->c2, (unsigned long long)add2);
        if (add1) atomicAdd(&a->c1, (unsigned long long)add1 << 32);
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < PADD_ENTRIES; e += PART_THREADS) {
      if (a_tag[e] == 0u) continue;
      JSlot* a = slots + (a_tag[e] - 1u);
      if (a_c2[e]) atomicAdd(&a->c2, (unsigned long long)a_c2[e]);
      if (a_c1[e]) atomicAdd(&a->c1, (unsigned long long)a_c1[e] << 32);
    }
    if (threadIdx.x == 0) pv.cur[8u * p] = 0u;
    __syncthreads();
  }
}
static_assert(PADD_ENTRIES == 512, "distinct_parts_kernel hashes junctions to 9 bits");

__device__ __forceinline__ fc_junction junction_from_slot(const JSlot& a, unsigned long long first_idx) {
  fc_junction o;
  o.chrom = (uint32_t)(a.kf >> (FIRST_BITS + 2)) & ((1u << CHROM_BITS) - 1u);
  o.start = (uint32_t)a.klo;
  o.end = (uint32_t)(a.klo >> 32);
  o.sk = ((uint32_t)(a.kf >> FIRST_BITS) & 3u) | (a.sig << 16);
  o.first_idx = first_idx;
  const uint32_t n_spanned = (uint32_t)a.c0, w8 = (uint32_t)(a.c0 >> 32);
  o.n_weighted = (double)w8 / 8.0;  // exact: every weight is k/8
  o.n_uniq_bridges = (double)(w8 - (uint32_t)a.c1) / 8.0;
  o.n_spanned = n_spanned;
  o.n_frags = n_spanned - (uint32_t)(a.c1 >> 32);
  // len(uniq)/2 with uniq = {read, revcomp(read)}: a palindromic read contributes one element, not two
  o.n_uniq = (2u * n_spanned - (uint32_t)a.c2) / 2u;
  const uint32_t lo = (uint32_t)a.ext, hi = (uint32_t)(a.ext >> 32);
  o.best_q_left = (int16_t)((int)(lo & 0xFFFFu) - 32768);
  o.best_q_right = (int16_t)((int)(lo >> 16) - 32768);
  o.min_n_hits = (uint16_t)(65535u - (hi >> 16));
  o.min_dist = (uint8_t)(255u - ((hi >> 8) & 0xFFu));
  o.min_ov = (uint8_t)(255u - (hi & 0xFFu));
  o.pad = 0;
  return o;
}

__device__ __forceinline__ void clear_slot(JSlot* a) {
  uint4* p = reinterpret_cast<uint4*>(a);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  p[0] = z;
  p[1] = z;
  p[2] = z;
  p[3] = z;
}

// Discovery order when the idx values of the records fill a known, dense range [idx_lo, idx_lo + range): the junction
// whose smallest idx is v is flagged at v - idx_lo; its rank is the number of flags before it.  Flags are counted per
// tile of RANK_TILE positions while they are set; the finish kernel adds the counts of the tiles before its own
// (given, or summed here when there are few tiles) to a block-wide scan of its tile.
constexpr int RANK_TILE = 1024;

__global__ void mark_first_kernel(unsigned int* __restrict__ ctr, unsigned int lcap, const unsigned int* __restrict__ list,
                                  const JSlot* __restrict__ slots, unsigned long long idx_lo, unsigned long long range,
                                  uint32_t* __restrict__ flag, uint32_t* __restrict__ tile_count) {
  const unsigned int n_alloc = min(ctr[FC_N_ALLOC], lcap);
  const unsigned int step = gridDim.x * blockDim.x;
  for (unsigned int j0 = blockIdx.x * blockDim.x; j0 < n_alloc; j0 += step) {  // warp-uniform trip count
    const unsigned int j = j0 + threadIdx.x;
    bool real = false;
    unsigned long long p = 0;
    if (j < n_alloc) {
      const unsigned int slot = list[j];
      p = FIRST_MAX - (slots[slot].kf & FIRST_MAX);  // (relative to idx_lo: the idx base of the accumulate pass)
      if (p < range) {
        real = true;
        flag[p] = slot + 1u;
      } else {
        atomicAdd(&ctr[FC_N_OVERFLOW], 1u);  // a record outside the declared idx range: the call goes to the sort-based path
      }
    }
    // early list entries are early discoveries: neighbouring lanes mostly hit the same tile, so count once per warp and tile
    const unsigned int tile = real ? (unsigned int)(p / RANK_TILE) : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xffffffffu, tile);
    if (real && (int)(threadIdx.x & 31) == __ffs((int)peers) - 1) atomicAdd(&tile_count[tile], (unsigned int)__popc(peers));
  }
}

constexpr int FINISH_THREADS = 256;
constexpr int FINISH_ITEMS = RANK_TILE / FINISH_THREADS;
static_assert(FINISH_ITEMS == 4, "finish_dense_kernel reads the flags of a thread as one uint4");

__global__ void __launch_bounds__(FINISH_THREADS) finish_dense_kernel(int64_t range, unsigned long long idx_lo,
                                                                       const uint32_t* __restrict__ flag,
                                                                       const uint32_t* __restrict__ tile_count,
                                                                       const uint32_t* __restrict__ tile_base,
                                                                       JSlot* __restrict__ slots, fc_junction* __restrict__ out,
                                                                       unsigned int* __restrict__ ctr,
                                                                       const unsigned long long* __restrict__ counters,
                                                                       unsigned long long* __restrict__ h_counters) {
  typedef cub::BlockScan<uint32_t, FINISH_THREADS> Scan;
  typedef cub::BlockReduce<uint32_t, FINISH_THREADS> Reduce;
  __shared__ union {
    typename Scan::TempStorage scan;
    typename Reduce::TempStorage reduce;
  } tmp;
  __shared__ uint32_t s_base;
  const bool last = blockIdx.x == gridDim.x - 1;
  const uint32_t mine = tile_count[blockIdx.x];
  if (mine == 0u && !last) return;  // (block-uniform) nothing was discovered in this stretch of the stream
  if (tile_base) {
    if (threadIdx.x == 0) s_base = tile_base[blockIdx.x];
  } else {
    uint32_t part = 0;
    for (unsigned int t = threadIdx.x; t < blockIdx.x; t += FINISH_THREADS) part += tile_count[t];
    const uint32_t before = Reduce(tmp.reduce).Sum(part);
    if (threadIdx.x == 0) s_base = before;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    // everything the host wants to know, written straight into its (mapped, pinned) memory: no copy to wait for
    ctr[FC_N_JUNC] = s_base + mine;
    __threadfence();
    for (int k = 0; k < 12; ++k) h_counters[k] = counters[k];
    h_counters[9] = (counters[9] & 0xFFFFFFFFull) | ((unsigned long long)(s_base + mine) << 32);
  }
  if (mine == 0u) return;
  const int64_t p0 = (int64_t)blockIdx.x * RANK_TILE + (int64_t)threadIdx.x * FINISH_ITEMS;
  uint32_t f[FINISH_ITEMS] = {0u, 0u, 0u, 0u};
  if (p0 + FINISH_ITEMS <= range) {
    const uint4 v = *reinterpret_cast<const uint4*>(flag + p0);
    f[0] = v.x;
    f[1] = v.y;
    f[2] = v.z;
    f[3] = v.w;
  } else {
    for (int k = 0; k < FINISH_ITEMS; ++k)
      if (p0 + k < range) f[k] = flag[p0 + k];
  }
  uint32_t local = 0;
  Scan(tmp.scan).ExclusiveSum((f[0] ? 1u : 0u) + (f[1] ? 1u : 0u) + (f[2] ? 1u : 0u) + (f[3] ? 1u : 0u), local);
  uint32_t o = s_base + local;
#pragma unroll
  for (int k = 0; k < FINISH_ITEMS; ++k) {
    if (!f[k]) continue;
    JSlot* ap = slots + (f[k] - 1u);
    const JSlot a = *ap;
    out[o++] = junction_from_slot(a, idx_lo + (unsigned long long)(p0 + k));
    clear_slot(ap);
  }
}

// records in arbitrary order (peer-to-peer emit, explicit idx): compact in any order, sorted by first idx afterwards
__global__ void finish_unordered_kernel(unsigned int* __restrict__ ctr, unsigned int lcap, const unsigned int* __restrict__ list,
                                        unsigned long long idx_base, JSlot* __restrict__ slots, fc_junction* __restrict__ out,
                                        uint64_t* __restrict__ order_key, uint32_t* __restrict__ order_val) {
  const unsigned int n_alloc = min(ctr[FC_N_ALLOC], lcap);
  for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_alloc; j += gridDim.x * blockDim.x) {
    JSlot* ap = slots + list[j];
    const JSlot a = *ap;
    const unsigned int pos = atomicAdd(&ctr[FC_N_JUNC], 1u);
    const unsigned long long first_idx = idx_base + (FIRST_MAX - (a.kf & FIRST_MAX));
    out[pos] = junction_from_slot(a, first_idx);
    order_key[pos] = first_idx;
    order_val[pos] = pos;
    clear_slot(ap);
  }
}

__global__ void finish_kernel(int64_t nj, int64_t n, const JAcc* __restrict__ acc, const fc_jrec* __restrict__ s,
                              fc_junction* __restrict__ out, uint64_t* __restrict__ order_key,
                              uint32_t* __restrict__ order_val) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  JAcc a = acc[j];
  const fc_jrec& r0 = s[a.seg_start];
  fc_junction o;
  o.chrom = r0.chrom;
  o.start = r0.start;
  o.end = r0.end;
  o.first_idx = a.first_idx;
  double w, b;
  uint32_t last_sk = r0.sk;
  if (a.cw_other == 0) {
    // every weight is k/8: any summation order gives this exact value
    w = (8.0 * a.cw[0] + 4.0 * a.cw[1] + 2.0 * a.cw[2] + 1.0 * a.cw[3]) / 8.0;
    b = (8.0 * a.cb[0] + 4.0 * a.cb[1] + 2.0 * a.cb[2] + 1.0 * a.cb[3]) / 8.0;
    last_sk = s[a.seg_start + a.n_spanned - 1].sk;
  } else {
    // replay in stream order, exactly as `self.n_weighted += weight` does (find_circ.py:544, 563)
    w = 0.0;
    b = 0.0;
    for (unsigned k = 0; k < a.n_spanned; ++k) {
      const fc_jrec& r = s[a.seg_start + k];
      double wt = 1.0 / (double)((r.sk >> 8) & 0xFFu);
      w += wt;
      if (r.q_left != 0 && r.q_right != 0) b += wt;
      last_sk = r.sk;
    }
  }
  o.sk = (r0.sk & 3u) | (last_sk & 0x0FFF0000u);  // signal of the LAST added splice (find_circ.py:528)
  o.n_weighted = w;
  o.n_uniq_bridges = b;
  o.n_spanned = a.n_spanned;
  o.n_frags = a.n_frags;
  // len(uniq)/2 with uniq = {read, revcomp(read)}: a palindromic read contributes one element, not two
  o.n_uniq = a.n_uniq - (a.n_pal + 1) / 2;
  o.best_q_left = (int16_t)a.max_ql;
  o.best_q_right = (int16_t)a.max_qr;
  o.min_n_hits = (uint16_t)a.min_nh;
  o.min_dist = (uint8_t)a.min_dist;
  o.min_ov = (uint8_t)a.min_ov;
  o.pad = 0;
  out[j] = o;
  order_key[j] = a.first_idx;
  order_val[j] = (uint32_t)j;
}

__global__ void gather_junctions_kernel(int64_t nj, const fc_junction* __restrict__ in, const uint32_t* __restrict__ perm,
                                        fc_junction* __restrict__ out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  const uint4* src = reinterpret_cast<const uint4*>(in + perm[j]);
  uint4* dst = reinterpret_cast<uint4*>(out + j);
  dst[0] = src[0];
  dst[1] = src[1];
  dst[2] = src[2];
  dst[3] = src[3];
}

__global__ void dest_rank_kernel(int64_t n, const fc_jrec* __restrict__ recs, int32_t n_ranks, uint64_t* __restrict__ key,
                                 uint32_t* __restrict__ iota) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const fc_jrec& r = recs[i];
  key[i] = fc_key_hash(r.chrom, r.start, r.end, r.sk, 0x5bd1e995ULL) % (uint64_t)n_ranks;
  iota[i] = (uint32_t)i;
}

__global__ void rank_hist_kernel(int64_t n, const uint64_t* __restrict__ key_sorted, unsigned long long* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // boundaries of the sorted destination array
  if (i == 0 || key_sorted[i] != key_sorted[i - 1]) {
    // start of rank key_sorted[i]; store start offsets, converted to counts on the host
    counts[key_sorted[i]] = (unsigned long long)i;
  }
}

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

int sort_pairs_u64_u32(fc_ctx* ctx, int64_t n, uint64_t* k_in, uint64_t* k_out, uint32_t* v_in, uint32_t* v_out,
                       int begin_bit, int end_bit, cudaStream_t st) {
  size_t tmp = 0;
  FC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, n, begin_bit, end_bit, st));
  FC_CUDA(ctx, ctx->agg.cub_tmp.reserve(tmp, st, false, 0));
  FC_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->agg.cub_tmp.p, tmp, k_in, k_out, v_in, v_out, n, begin_bit, end_bit, st));
  ctx->launches += 1 + (end_bit - begin_bit + 7) / 8;
  return FC_OK;
}
int scan_u32(fc_ctx* ctx, int64_t n, const uint32_t* in, uint32_t* out, bool inclusive, cudaStream_t st) {
  size_t tmp = 0;
  if (inclusive) {
    FC_CUDA(ctx, cub::DeviceScan::InclusiveSum(nullptr, tmp, in, out, n, st));
    FC_CUDA(ctx, ctx->agg.cub_tmp.reserve(tmp, st, false, 0));
    FC_CUDA(ctx, cub::DeviceScan::InclusiveSum(ctx->agg.cub_tmp.p, tmp, in, out, n, st));
  } else {
    FC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, st));
    FC_CUDA(ctx, ctx->agg.cub_tmp.reserve(tmp, st, false, 0));
    FC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->agg.cub_tmp.p, tmp, in, out, n, st));
  }
  ctx->launches += 2;
  return FC_OK;
}

// counters: words [0, FC_RESET_WORDS) are zeroed by fc_agg_reset*; word FC_BARRIER_WORD counts barrier arrivals for the
// whole life of the context; word FC_BARRIER_TIMEOUT_WORD (copied to the host by fc_agg_finalize) flags a barrier that gave up
constexpr int FC_RESET_WORDS = 16;
constexpr int FC_BARRIER_WORD = 32;
constexpr int FC_BARRIER_TIMEOUT_WORD = 5;

int ensure_counters(fc_ctx* ctx, cudaStream_t st) {
  if (!ctx->agg.counters.p) {
    FC_CUDA(ctx, ctx->agg.counters.reserve(fc::FC_CNT_WORDS * sizeof(unsigned long long), st, false, 0));
    FC_CUDA(ctx, cudaMemsetAsync(ctx->agg.counters.p, 0, fc::FC_CNT_WORDS * sizeof(unsigned long long), st));
  }
  if (!ctx->agg.h_pinned) FC_CUDA(ctx, cudaHostAlloc((void**)&ctx->agg.h_pinned, 64 * sizeof(unsigned long long), cudaHostAllocMapped));
  return FC_OK;
}

int sync_n_recs(fc_ctx* ctx, cudaStream_t st) {
  if (!ctx->agg.counters.p) {
    ctx->agg.n_recs = 0;
    return FC_OK;
  }
  if (ctx->agg.n_exact) return FC_OK;  // the host already knows the exact count
  unsigned long long v = 0;
  FC_CUDA(ctx, cudaMemcpyAsync(&v, ctx->agg.counters.p, sizeof(v), cudaMemcpyDeviceToHost, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->agg.n_recs = (int64_t)v;
  ctx->agg.n_exact = true;
  return FC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ fc_dbuf
cudaError_t fc_dbuf::reserve(size_t bytes, cudaStream_t st, bool keep, size_t used) {
  if (bytes <= cap) return cudaSuccess;
  size_t ncap = cap ? cap : 4096;
  while (ncap < bytes) ncap = ncap + ncap / 2 + 4096;
  void* np = nullptr;
  cudaError_t e = cudaMalloc(&np, ncap);
  if (e != cudaSuccess) return e;
  if (p) {
    if (keep && used) {
      e = cudaMemcpyAsync(np, p, used, cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return e;
    }
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    cudaFree(p);
  }
  p = np;
  cap = ncap;
  return cudaSuccess;
}
void fc_dbuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

// ------------------------------------------------------------------------------------------------ C ABI
static int clear_sets_early(fc_ctx* ctx, cudaStream_t st);

extern "C" int fc_agg_reset(fc_ctx* ctx) {
  if (!ctx) return FC_E_ARG;
  ctx->agg.n_recs = 0;
  ctx->agg.n_exact = true;
  ctx->agg.unordered = false;
  ctx->agg.n_junc = -1;
  ctx->agg.max_idx = 0;
  ctx->agg.idx_lo = ~0ull;
  ctx->agg.range_declared = false;
  if (ctx->agg.counters.p) FC_CUDA(ctx, cudaMemset(ctx->agg.counters.p, 0, FC_RESET_WORDS * sizeof(unsigned long long)));
  return clear_sets_early(ctx, ctx->own_stream);
}

extern "C" int fc_agg_reset_async(fc_ctx* ctx, void* stream) {
  if (!ctx) return FC_E_ARG;
  fc_agg& a = ctx->agg;
  int rc = ensure_counters(ctx, (cudaStream_t)stream);
  if (rc) return rc;
  a.n_recs = 0;
  a.n_exact = true;
  a.unordered = false;
  a.n_junc = -1;
  a.max_idx = 0;
  a.idx_lo = ~0ull;
  a.range_declared = false;
  FC_CUDA(ctx, cudaMemsetAsync(a.counters.p, 0, FC_RESET_WORDS * sizeof(unsigned long long), (cudaStream_t)stream));
  return clear_sets_early(ctx, (cudaStream_t)stream);
}

int fc_agg_reserve_records(fc_ctx* ctx, int64_t extra, cudaStream_t st) {
  fc_agg& a = ctx->agg;
  // the counters exist (and their initial memset is queued on `st`) before the caller forks work onto other streams
  int rc = ensure_counters(ctx, st);
  if (rc) return rc;
  FC_CUDA(ctx, a.recs.reserve((size_t)(a.n_recs + extra) * sizeof(fc_jrec), st, true, (size_t)a.n_recs * sizeof(fc_jrec)));
  return FC_OK;
}

// bookkeeping around an emit (the stand-alone kernel below or the scan kernel that emits on the way, scan.cu)
int fc_agg_emit_begin(fc_ctx* ctx, int64_t n, cudaStream_t st, fc::EmitArgs* e) {
  if (n >= (1ll << 32)) return fc_fail(ctx, FC_E_ARG, "batch too large");
  fc_agg& a = ctx->agg;
  if (a.p2p_enabled)
    return fc_fail(ctx, FC_E_STATE, "this context is connected to peers and reduces what they send: record with the peer emit (fc_scan_emit_batch, "
                                    "fc_scan_emit_p2p, fc_agg_emit_p2p), or aggregate locally in another context");
  int rc = ensure_counters(ctx, st);
  if (rc) return rc;
  // upper bound of the record count so far (the exact count lives on the device)
  FC_CUDA(ctx, a.recs.reserve((size_t)(a.n_recs + n) * sizeof(fc_jrec), st, true, (size_t)a.n_recs * sizeof(fc_jrec)));
  e->n_recs = (unsigned long long*)a.counters.p;
  e->recs = (fc_jrec*)a.recs.p;
  return FC_OK;
}

void fc_agg_emit_end(fc_ctx* ctx, int64_t n, uint64_t idx_base, bool explicit_idx) {
  fc_agg& a = ctx->agg;
  a.n_recs += n;  // upper bound until the next sync
  a.n_exact = false;
  a.n_junc = -1;
  a.unordered = true;  // slots are claimed per CTA: consumers that need stream order restore it from idx
  if (a.range_declared) {
    // the caller has declared the idx range of everything that arrives
  } else if (explicit_idx) {
    a.max_idx = ~0ull;  // explicit positions: range unknown
  } else if (a.max_idx != ~0ull) {
    if (idx_base + (uint64_t)n > a.max_idx) a.max_idx = idx_base + (uint64_t)n;
    if (idx_base < a.idx_lo) a.idx_lo = idx_base;
  }
}

static int agg_emit_impl(fc_ctx* ctx, int64_t n, const fc_hit* d_hits, const int32_t* d_chrom, const uint8_t* d_flags,
                           const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b,
                           const uint64_t* d_read_hash, const uint64_t* d_qname_hash, const uint8_t* d_mask,
                           uint64_t idx_base, const uint64_t* d_idx, void* stream) {
  if (!ctx || n < 0) return FC_E_ARG;
  if (n == 0) return FC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  fc::EmitArgs e{d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, d_idx, nullptr, nullptr};
  int rc = fc_agg_emit_begin(ctx, n, st, &e);
  if (rc) return rc;
  emit_kernel<<<nblk(n, 256), 256, 0, st>>>(n, d_hits, d_mask, d_chrom, d_flags, e);
  FC_LAUNCH_CHECK(ctx);
  fc_agg_emit_end(ctx, n, idx_base, d_idx != nullptr);
  return FC_OK;
}

extern "C" int fc_agg_emit(fc_ctx* ctx, int64_t n, const fc_hit* d_hits, const int32_t* d_chrom, const uint8_t* d_flags,
                           const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b,
                           const uint64_t* d_read_hash, const uint64_t* d_qname_hash, const uint8_t* d_mask,
                           uint64_t idx_base, void* stream) {
  return agg_emit_impl(ctx, n, d_hits, d_chrom, d_flags, d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, d_mask, idx_base,
                       nullptr, stream);
}

extern "C" int fc_agg_emit_idx(fc_ctx* ctx, int64_t n, const fc_hit* d_hits, const int32_t* d_chrom, const uint8_t* d_flags,
                               const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b,
                               const uint64_t* d_read_hash, const uint64_t* d_qname_hash, const uint8_t* d_mask,
                               const uint64_t* d_idx, void* stream) {
  if (!d_idx) return FC_E_ARG;
  return agg_emit_impl(ctx, n, d_hits, d_chrom, d_flags, d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, d_mask, 0, d_idx,
                       stream);
}

extern "C" int fc_agg_append(fc_ctx* ctx, int64_t n, const fc_jrec* d_recs, void* stream) {
  if (!ctx || n < 0) return FC_E_ARG;
  if (n == 0) return FC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  fc_agg& a = ctx->agg;
  if (a.p2p_enabled) return fc_fail(ctx, FC_E_STATE, "fc_agg_append on a context that is connected to peers");
  int rc = ensure_counters(ctx, st);
  if (rc) return rc;
  rc = sync_n_recs(ctx, st);
  if (rc) return rc;
  FC_CUDA(ctx, a.recs.reserve((size_t)(a.n_recs + n) * sizeof(fc_jrec), st, true, (size_t)a.n_recs * sizeof(fc_jrec)));
  FC_CUDA(ctx, cudaMemcpyAsync((fc_jrec*)a.recs.p + a.n_recs, d_recs, (size_t)n * sizeof(fc_jrec), cudaMemcpyDefault, st));
  a.n_recs += n;
  if (!a.range_declared) a.max_idx = ~0ull;  // records built elsewhere: their idx range and order are unknown
  a.unordered = true;
  unsigned long long v = (unsigned long long)a.n_recs;
  FC_CUDA(ctx, cudaMemcpyAsync(a.counters.p, &v, sizeof(v), cudaMemcpyHostToDevice, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  a.n_junc = -1;
  return FC_OK;
}

__global__ void set_count_kernel(unsigned long long* c, unsigned long long v) { *c = v; }

// replace the record buffer by n records given on the device (the receive side of the multi-GPU exchange);
// asynchronous on `stream`, no host synchronisation
extern "C" int fc_agg_replace(fc_ctx* ctx, int64_t n, const fc_jrec* d_recs, void* stream) {
  if (!ctx || n < 0) return FC_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  fc_agg& a = ctx->agg;
  int rc = ensure_counters(ctx, st);
  if (rc) return rc;
  FC_CUDA(ctx, a.recs.reserve((size_t)(n > 0 ? n : 1) * sizeof(fc_jrec), st, false, 0));
  if (n > 0) FC_CUDA(ctx, cudaMemcpyAsync(a.recs.p, d_recs, (size_t)n * sizeof(fc_jrec), cudaMemcpyDeviceToDevice, st));
  set_count_kernel<<<1, 1, 0, st>>>((unsigned long long*)a.counters.p, (unsigned long long)n);
  FC_LAUNCH_CHECK(ctx);
  a.n_recs = n;
  a.n_exact = true;
  if (!a.range_declared) a.max_idx = ~0ull;
  a.unordered = true;
  a.n_junc = -1;
  return FC_OK;
}

extern "C" int fc_agg_append_host(fc_ctx* ctx, int64_t n, const fc_jrec* h_recs) {
  if (!ctx) return FC_E_ARG;
  return fc_agg_append(ctx, n, h_recs, ctx->own_stream);
}

extern "C" int64_t fc_agg_n_records(fc_ctx* ctx) {
  if (!ctx) return FC_E_ARG;
  int rc = sync_n_recs(ctx, ctx->own_stream);
  if (rc) return rc;
  // the record buffer may be in use on another stream: make sure everything has landed
  cudaDeviceSynchronize();
  rc = sync_n_recs(ctx, ctx->own_stream);
  if (rc) return rc;
  return ctx->agg.n_recs;
}

extern "C" const fc_jrec* fc_agg_records(fc_ctx* ctx) { return ctx ? (const fc_jrec*)ctx->agg.recs.p : nullptr; }

extern "C" int fc_agg_partition(fc_ctx* ctx, int32_t n_ranks, fc_jrec* d_out, int64_t* h_counts, void* stream) {
  if (!ctx || n_ranks <= 0 || !h_counts) return FC_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  fc_agg& a = ctx->agg;
  int rc = sync_n_recs(ctx, st);
  if (rc) return rc;
  int64_t n = a.n_recs;
  for (int r = 0; r < n_ranks; ++r) h_counts[r] = 0;
  if (n == 0) return FC_OK;
  if (!d_out) return FC_E_ARG;
  FC_CUDA(ctx, a.scratch[0].reserve((size_t)n * 8, st, false, 0));
  FC_CUDA(ctx, a.scratch[1].reserve((size_t)n * 8, st, false, 0));
  FC_CUDA(ctx, a.scratch[2].reserve((size_t)n * 4, st, false, 0));
  FC_CUDA(ctx, a.scratch[3].reserve((size_t)n * 4, st, false, 0));
  FC_CUDA(ctx, a.scratch[4].reserve((size_t)(n_ranks + 1) * 8, st, false, 0));
  uint64_t* key = (uint64_t*)a.scratch[0].p;
  uint64_t* key_s = (uint64_t*)a.scratch[1].p;
  uint32_t* iota = (uint32_t*)a.scratch[2].p;
  uint32_t* perm = (uint32_t*)a.scratch[3].p;
  unsigned long long* starts = (unsigned long long*)a.scratch[4].p;
  dest_rank_kernel<<<nblk(n, 256), 256, 0, st>>>(n, (const fc_jrec*)a.recs.p, n_ranks, key, iota);
  FC_LAUNCH_CHECK(ctx);
  int bits = 1;
  while ((1 << bits) < n_ranks) bits++;
  rc = sort_pairs_u64_u32(ctx, n, key, key_s, iota, perm, 0, bits, st);  // stable: stream order kept per destination
  if (rc) return rc;
  gather_kernel<<<nblk(n, 256), 256, 0, st>>>(n, (const fc_jrec*)a.recs.p, perm, d_out);
  FC_LAUNCH_CHECK(ctx);
  FC_CUDA(ctx, cudaMemsetAsync(starts, 0xFF, (size_t)(n_ranks + 1) * 8, st));
  rank_hist_kernel<<<nblk(n, 256), 256, 0, st>>>(n, key_s, starts);
  FC_LAUNCH_CHECK(ctx);
  std::vector<unsigned long long> h(n_ranks + 1);
  FC_CUDA(ctx, cudaMemcpyAsync(h.data(), starts, (size_t)(n_ranks + 1) * 8, cudaMemcpyDeviceToHost, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  h[n_ranks] = (unsigned long long)n;
  // ranks that received nothing keep the 0xFF.. marker: their start is the next valid start
  for (int r = n_ranks - 1; r >= 0; --r)
    if (h[r] == ~0ull) h[r] = h[r + 1];
  for (int r = 0; r < n_ranks; ++r) h_counts[r] = (int64_t)(h[r + 1] - h[r]);
  return FC_OK;
}

// reserve a table that the fused path keeps clean between calls; a (re)allocation hands out fresh memory: zero it
static int reserve_clean(fc_ctx* ctx, fc_dbuf& b, size_t bytes, cudaStream_t st) {
  if (bytes <= b.cap) return FC_OK;
  FC_CUDA(ctx, b.reserve(bytes, st, false, 0));
  FC_CUDA(ctx, cudaMemsetAsync(b.p, 0, b.cap, st));
  return FC_OK;
}

// The distinct set has to be all-zero when the accumulate kernel starts.  fc_agg_reset* already knows that, so it clears
// the part the last call dirtied on a side stream: the memset then runs beside the scan kernel of the next batch
// instead of in front of the accumulate kernel.
static int clear_sets_early(fc_ctx* ctx, cudaStream_t st) {
  fc_agg& a = ctx->agg;
  if (!a.f_sets.p || a.sets_clean || a.sets_used == 0) return FC_OK;
  if (!a.side) {
    FC_CUDA(ctx, cudaStreamCreateWithFlags(&a.side, cudaStreamNonBlocking));
    FC_CUDA(ctx, cudaEventCreateWithFlags(&a.ev_side, cudaEventDisableTiming));
    FC_CUDA(ctx, cudaEventCreateWithFlags(&a.ev_main, cudaEventDisableTiming));
  }
  FC_CUDA(ctx, cudaEventRecord(a.ev_main, st));
  FC_CUDA(ctx, cudaStreamWaitEvent(a.side, a.ev_main, 0));
  FC_CUDA(ctx, cudaMemsetAsync(a.f_sets.p, 0, a.sets_used, a.side));
  FC_CUDA(ctx, cudaEventRecord(a.ev_side, a.side));
  a.sets_clean = true;
  return FC_OK;
}

// per-stage device times of the sort-free path (CUDA events on the caller's stream): fc_agg_set_timing() keeps them
// for fc_agg_get_timing(), FC_AGG_TIMING=1 also prints them on stderr
struct StageTimer {
  bool on, print;
  fc_ctx* ctx;
  cudaStream_t st;
  int n = 0;
  cudaEvent_t ev[12];
  const char* name[12];
  StageTimer(fc_ctx* c, cudaStream_t s) : ctx(c), st(s) {
    static int flag = -1;
    if (flag < 0) {
      const char* e = getenv("FC_AGG_TIMING");
      flag = (e && e[0] == '1') ? 1 : 0;
    }
    print = flag == 1;
    on = print || c->agg.timing;
  }
  void mark(const char* what) {
    if (!on || n >= 12) return;
    cudaEventCreate(&ev[n]);
    cudaEventRecord(ev[n], st);
    name[n++] = what;
  }
  void report() {
    if (!on) return;
    for (int k = 1; k < n; ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[k - 1], ev[k]);
      if (k - 1 < 8) ctx->agg.stage_us[k - 1] = ms * 1000.f;
      if (print) fprintf(stderr, "%s%s %.1f us", k == 1 ? "[fc_agg_finalize] " : ", ", name[k], ms * 1000.f);
    }
    if (print) fprintf(stderr, "\n");
    for (int k = 0; k < n; ++k) cudaEventDestroy(ev[k]);
    n = 0;
  }
};

// sort-free path over the `ub` (upper bound; the exact count is on the device) records of the context; returns -100
// when the input needs the sort-based path (a weight denominator that is not 1, 2, 4 or 8; too many records)
static int64_t finalize_fused(fc_ctx* ctx, int64_t ub, cudaStream_t st, const RecSrc& src, bool global_set = false, bool full_table = false) {
  fc_agg& a = ctx->agg;
  if (ub >= FUSED_MAX_RECORDS) return -100;
  // junction table: one 64-byte slot per junction, at most one junction per record (load <= 0.8); distinct set: up to two
  // entries per record (load <= 2/3 when every read and every name is new)
  unsigned long long kcap = 1024, scap = 2048;
  while (4ull * kcap < 5ull * (unsigned long long)ub) kcap <<= 1;
  while (scap < 3ull * (unsigned long long)ub) scap <<= 1;
  // The slots of a call lie all over the table, and a table of gigabytes costs every access a TLB miss on top (measured at
  // config 3, 426 k junctions: 2.32 ms with the 4-GB table the record count asks for, 1.68 ms with 128 MB).  So a context
  // that has reduced a batch before sizes the table after that batch's junction count (four times it: load 0.25 when the
  // next batch is alike); a batch with so many more junctions that records find no place (SLOT_MAX_PROBES) is detected and
  // reduced again with the full table.
  int max_probes = 0x7fffffff;  // the table the record count asks for has room for every junction: probe until found
  if (!full_table && a.nj_hint >= 0) {
    unsigned long long want = 4ull * (unsigned long long)a.nj_hint + 32768ull, k = 1024;
    if (const char* e = getenv("FC_AGG_TABLE_HINT")) want = (unsigned long long)atoll(e);  // (tests: a table that is too small)
    while (k < want) k <<= 1;
    if (k < kcap) {
      kcap = k;
      max_probes = SLOT_MAX_PROBES;
    }
  }
  const unsigned int lcap = (unsigned int)ub + 1024u;  // list of the occupied slots (one entry per junction)
  int rc;
  StageTimer tm(ctx, st);
  tm.mark("start");
  if (a.side) FC_CUDA(ctx, cudaStreamWaitEvent(st, a.ev_side, 0));  // a pending early clear
  if (a.f_dirty) {  // an earlier call failed half-way: start from clean tables
    if (a.side) FC_CUDA(ctx, cudaStreamSynchronize(a.side));
    a.sets_clean = false;
    a.f_keys.release();
    a.f_sets.release();
    a.f_pcur.release();  // (the partition cursors as well: a kernel that did not finish has left counts behind)
    a.f_dirty = false;
  }
  if ((rc = reserve_clean(ctx, a.f_keys, (size_t)kcap * sizeof(JSlot), st))) return rc;
  FC_CUDA(ctx, a.f_acc.reserve((size_t)lcap * sizeof(unsigned int), st, false, 0));
  // distinct counts: one global set while it fits L2, else partitions + shared-memory sets (FC_AGG_SETS=global|part forces one)
  bool part = (size_t)scap * 16 > ((size_t)64 << 20);
  if (const char* e = getenv("FC_AGG_SETS")) part = e[0] == 'p' ? true : (e[0] == 'g' ? false : part);
  if (global_set) part = false;
  PartView pv{nullptr, nullptr, 0u, 0u};
  if (part) {
    pv.n_parts = (unsigned int)((ub + PSET_TARGET - 1) / PSET_TARGET);
    if (pv.n_parts == 0u) pv.n_parts = 1u;
    pv.pcap = 4u * PSET_TARGET + 2048u;  // (twice the mean when every name goes through the set too, and as much again for repeats)
    if (const char* e = getenv("FC_AGG_PART_CAP")) pv.pcap = (unsigned int)atoi(e) > 0 ? (unsigned int)atoi(e) : pv.pcap;  // (tests: force the way back)
    pv.pcap = (pv.pcap + PART_CHUNK - 1u) & ~(PART_CHUNK - 1u);
    if ((rc = reserve_clean(ctx, a.f_pcur, (size_t)pv.n_parts * 32, st))) return rc;
    FC_CUDA(ctx, a.f_part.reserve((size_t)pv.n_parts * pv.pcap * sizeof(uint4), st, false, 0));
    pv.ent = (uint4*)a.f_part.p;
    pv.cur = (unsigned int*)a.f_pcur.p;
  } else {
    if (!(a.sets_clean && (size_t)scap * 16 <= a.sets_used && (size_t)scap * 16 <= a.f_sets.cap)) {
      FC_CUDA(ctx, a.f_sets.reserve((size_t)scap * 16, st, false, 0));
      FC_CUDA(ctx, cudaMemsetAsync(a.f_sets.p, 0, (size_t)scap * 16, st));
    }
    a.sets_clean = false;
    a.sets_used = (size_t)scap * 16 > a.sets_used ? (size_t)scap * 16 : a.sets_used;
  }
  FC_CUDA(ctx, a.junctions.reserve((size_t)ub * sizeof(fc_junction), st, false, 0));
  unsigned long long* counters = (unsigned long long*)a.counters.p;
  unsigned int* ctr = (unsigned int*)(counters + 8);
  FC_CUDA(ctx, cudaMemsetAsync(ctr, 0, FC_N_CTR * sizeof(unsigned int), st));
  // discovery rank: flags over the idx range when it is known and about as large as the record count, else a sort
  const bool dense = a.max_idx != ~0ull && a.idx_lo != ~0ull && a.max_idx > a.idx_lo &&
                     a.max_idx - a.idx_lo <= 4ull * (unsigned long long)ub + (1ull << 20);
  const int64_t range = dense ? (int64_t)(a.max_idx - a.idx_lo) : 0;
  const unsigned long long idx_base = a.idx_lo != ~0ull ? a.idx_lo : 0ull;  // record positions are kept relative to it (JSlot.kf)
  const int64_t n_tiles = (range + RANK_TILE - 1) / RANK_TILE;
  a.f_dirty = true;  // until the finish kernel has run
  uint32_t* flag = nullptr;
  uint32_t* tile_count = nullptr;
  if (dense) {
    FC_CUDA(ctx, a.scratch[0].reserve((size_t)range * 4 + 16, st, false, 0));  // (cleared 16 bytes at a time)
    FC_CUDA(ctx, a.scratch[1].reserve((size_t)n_tiles * 8, st, false, 0));
    flag = (uint32_t*)a.scratch[0].p;
    tile_count = (uint32_t*)a.scratch[1].p;
  }
  tm.mark("clear");
  {
    // persistent CTAs: two per SM, each walks chunks of up to ACC_MAX_TILES tiles of ACC_THREADS records; small inputs get
    // shorter chunks so that every SM has work
    const int64_t tiles = (ub + ACC_THREADS - 1) / ACC_THREADS;
    const int64_t ctas = (int64_t)ctx->sm_count * FC_ACC_MIN_CTAS;
    int chunk_tiles = (int)((tiles + 2 * ctas - 1) / (2 * ctas));
    chunk_tiles = chunk_tiles < 1 ? 1 : (chunk_tiles > ACC_MAX_TILES ? ACC_MAX_TILES : chunk_tiles);
    const int64_t chunks = (tiles + chunk_tiles - 1) / chunk_tiles;
    const unsigned grid = (unsigned)(chunks < ctas ? chunks : ctas);
    // the L2 prefetch of the set slots pays while the set fits L2 (-7 % at 1 M records) and costs 20 % when it does not
    const int prefetch = (size_t)scap * 16 <= ((size_t)64 << 20) ? 1 : 0;
    if (part) FC_CUDA(ctx, cudaFuncSetAttribute(fused_accumulate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RECENT_SETS * 8));
    if (part)
      fused_accumulate_kernel<1><<<grid, ACC_THREADS, 2 * RECENT_SETS * 8, st>>>(src, chunk_tiles, 0, max_probes, idx_base, (JSlot*)a.f_keys.p, kcap - 1, nullptr, 0ull, pv,
                                                               (unsigned int*)a.f_acc.p, lcap, ctr, (uint4*)flag, (range + 3) / 4,
                                                               tile_count, n_tiles);
    else
      fused_accumulate_kernel<0><<<grid, ACC_THREADS, 0, st>>>(src, chunk_tiles, prefetch, max_probes, idx_base, (JSlot*)a.f_keys.p, kcap - 1,
                                                               (U128*)a.f_sets.p, scap - 1, pv, (unsigned int*)a.f_acc.p, lcap, ctr,
                                                               (uint4*)flag, (range + 3) / 4, tile_count, n_tiles);
  }
  FC_LAUNCH_CHECK(ctx);
  tm.mark("accumulate");
  if (part) {
    FC_CUDA(ctx, cudaFuncSetAttribute(distinct_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PSET_ENTRIES * 8));
    const unsigned want = (unsigned)ctx->sm_count * 3u;
    distinct_parts_kernel<<<pv.n_parts < want ? pv.n_parts : want, PART_THREADS, PSET_ENTRIES * 8, st>>>(pv, (JSlot*)a.f_keys.p, ctr);
    FC_LAUNCH_CHECK(ctx);
    ctx->launches += 1;
  }
  tm.mark("distinct");
  const unsigned sweep_blocks = (unsigned)ctx->sm_count * 8u;
  uint64_t* kA = nullptr;
  uint64_t* kB = nullptr;
  uint32_t* vA = nullptr;
  fc_junction* tmpj = nullptr;
  if (dense) {
    mark_first_kernel<<<sweep_blocks, 256, 0, st>>>(ctr, lcap, (const unsigned int*)a.f_acc.p, (const JSlot*)a.f_keys.p, a.idx_lo,
                                                    (unsigned long long)range, flag, tile_count);
    FC_LAUNCH_CHECK(ctx);
    uint32_t* tile_base = nullptr;
    if (n_tiles > 4096) {  // many tiles: scan the counts once instead of summing them in every block
      tile_base = tile_count + n_tiles;
      size_t tmp = 0;
      FC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, tile_count, tile_base, n_tiles, st));
      FC_CUDA(ctx, a.cub_tmp.reserve(tmp, st, false, 0));
      FC_CUDA(ctx, cub::DeviceScan::ExclusiveSum(a.cub_tmp.p, tmp, tile_count, tile_base, n_tiles, st));
      ctx->launches += 2;
    }
    tm.mark("mark");
    finish_dense_kernel<<<(unsigned)n_tiles, FINISH_THREADS, 0, st>>>(range, a.idx_lo, flag, tile_count, tile_base,
                                                                       (JSlot*)a.f_keys.p, (fc_junction*)a.junctions.p, ctr, counters,
                                                                       a.h_pinned);
    FC_LAUNCH_CHECK(ctx);
  } else {
    FC_CUDA(ctx, a.scratch[5].reserve((size_t)ub * sizeof(fc_junction), st, false, 0));
    FC_CUDA(ctx, a.scratch[3].reserve((size_t)ub * 8, st, false, 0));
    FC_CUDA(ctx, a.scratch[4].reserve((size_t)ub * 8, st, false, 0));
    FC_CUDA(ctx, a.scratch[6].reserve((size_t)ub * 8, st, false, 0));
    tmpj = (fc_junction*)a.scratch[5].p;
    kA = (uint64_t*)a.scratch[3].p;
    kB = (uint64_t*)a.scratch[4].p;
    vA = (uint32_t*)a.scratch[6].p;
    finish_unordered_kernel<<<sweep_blocks, 256, 0, st>>>(ctr, lcap, (const unsigned int*)a.f_acc.p, idx_base, (JSlot*)a.f_keys.p, tmpj, kA, vA);
    FC_LAUNCH_CHECK(ctx);
  }
  // one round trip: exact record count, junction count, fallback conditions, peer-to-peer overflow
  tm.mark("finish");
  unsigned long long* h = a.h_pinned;  // pinned + mapped; the dense finish kernel has already written it
  if (!dense) FC_CUDA(ctx, cudaMemcpyAsync(h, counters, 12 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  tm.mark("copy");
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  tm.report();
  a.f_dirty = false;
  a.n_recs = (int64_t)h[0];
  a.n_exact = true;
  if (a.p2p_enabled && h[FC_BARRIER_TIMEOUT_WORD])
    return fc_fail(ctx, FC_E_STATE, "a peer-memory barrier timed out (a peer rank did not arrive in %.0f s): the results of this step are invalid",
                   a.p2p_timeout_s);
  if (a.p2p_enabled && h[4])
    return fc_fail(ctx, FC_E_NOMEM, "peer-to-peer record buffer overflow (%llu records dropped): raise the capacity", h[4]);
  if (h[6])
    return fc_fail(ctx, FC_E_ARG, "%llu records came without a read-name hash and without fragment fields in their descriptors", h[6]);
  const unsigned int n_other = (unsigned int)(h[8] >> 32), n_overflow = (unsigned int)h[9];
  const int64_t nj = (int64_t)(h[9] >> 32);
  if (n_overflow) a.f_dirty = true;  // ids ran out, or a record lies outside a declared idx range: its accumulator was not consumed
  if ((unsigned int)h[11] && max_probes == 0x7fffffff)
    return fc_fail(ctx, FC_E_STATE, "junction table: no free slot in a table with room for every record (internal error)");
  if ((unsigned int)h[11]) {
    // the junction table was too small for this batch: everything is clean again (the lanes without a slot touched
    // nothing, the finish pass has consumed what the others did): once more with the table the record count asks for
    if (tm.print) fprintf(stderr, "[fc_agg_finalize] junction table of %llu slots too small (%u records without a slot): full table\n", kcap, (unsigned)h[11]);
    a.nj_hint = -1;
    return finalize_fused(ctx, ub, st, src, global_set, true);
  }
  if (n_overflow) {
    a.range_declared = false;
    a.max_idx = ~0ull;
  }
  if (n_other || n_overflow) return -100;
  if (part && h[10]) {
    if (tm.print)
      fprintf(stderr, "[fc_agg_finalize] partitioned distinct counts gave up (%u entries beyond a partition, %u beyond a set): global set\n",
              (unsigned)h[10], (unsigned)(h[10] >> 32));
    // a partition or its shared-memory set ran full (copies of one element beyond what the caches absorb, or a hash that
    // does not spread): every table is clean again, the call is repeated with the global set
    return finalize_fused(ctx, ub, st, src, true, full_table);
  }
  if (!dense && nj > 0) {
    uint32_t* vB = vA + ub;
    rc = sort_pairs_u64_u32(ctx, nj, kA, kB, vA, vB, 0, 64, st);
    if (rc) return rc;
    gather_junctions_kernel<<<nblk(nj, 256), 256, 0, st>>>(nj, tmpj, vB, (fc_junction*)a.junctions.p);
    FC_LAUNCH_CHECK(ctx);
  }
  a.n_junc = nj;
  a.nj_hint = nj;
  return nj;
}

// multi-GPU: the slice counts the source ranks have published for the current step (host copy), clamped to the capacity
static int p2p_slice_counts(fc_ctx* ctx, cudaStream_t st, unsigned long long* out /* 8 */, int64_t* total) {
  fc_agg& a = ctx->agg;
  const int parity = (int)((a.barrier_epoch - 1) & 1ull);
  unsigned long long blk[64];
  FC_CUDA(ctx, cudaMemcpyAsync(blk, a.counters.p, sizeof(blk), cudaMemcpyDeviceToHost, st));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  if (blk[FC_BARRIER_TIMEOUT_WORD])
    return fc_fail(ctx, FC_E_STATE, "a peer-memory barrier timed out (a peer rank did not arrive in %.0f s): the results of this step are invalid",
                   a.p2p_timeout_s);
  if (blk[4])
    return fc_fail(ctx, FC_E_NOMEM, "peer-to-peer record buffer overflow (%llu records of this rank dropped): raise the capacity", blk[4]);
  if (blk[6])
    return fc_fail(ctx, FC_E_ARG, "%llu records came without a read-name hash and without fragment fields in their descriptors", blk[6]);
  const unsigned long long* h = blk + fc::FC_CNT_SLICE + 8 * parity;
  *total = 0;
  for (int r = 0; r < 8; ++r) {
    out[r] = 0;
    if (r >= a.p2p_world) continue;
    if (h[r] > (unsigned long long)a.p2p_slice_cap)
      return fc_fail(ctx, FC_E_NOMEM, "peer-to-peer record buffer overflow (rank %d sent %llu records into a slice of %lld): raise the capacity",
                     r, h[r], (long long)a.p2p_slice_cap);
    out[r] = h[r];
    *total += (int64_t)h[r];
  }
  return FC_OK;
}

extern "C" int64_t fc_agg_finalize(fc_ctx* ctx, void* stream) {
  if (!ctx) return FC_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  fc_agg& a = ctx->agg;
  if (!a.counters.p || (!a.p2p_enabled && a.n_exact && a.n_recs == 0)) {
    a.n_recs = 0;
    a.n_junc = 0;
    return 0;
  }
  // FC_AGG_MODE=sort forces the sort-based path (tests compare the two)
  const char* mode_env = getenv("FC_AGG_MODE");
  const bool sort_mode = mode_env && mode_env[0] == 's';
  fc_jrec* rbuf = (fc_jrec*)a.recs.p;  // where the sort-based path finds (and re-orders) the records
  if (a.p2p_enabled) {
    // records sit in one slice per source rank in the half of the buffer that belongs to the step the last barrier closed
    if (a.p2p_open) return fc_fail(ctx, FC_E_STATE, "fc_agg_finalize before the fc_p2p_barrier that ends the step");
    if (a.barrier_epoch == 0) {
      a.n_recs = 0;
      a.n_junc = 0;
      return 0;
    }
    unsigned long long cnt[8];
    int64_t total = 0;
    int rc = p2p_slice_counts(ctx, st, cnt, &total);
    if (rc) return rc;
    a.n_recs = total;
    a.n_exact = true;
    if (total == 0) {
      a.n_junc = 0;
      return 0;
    }
    if (total >= (1ll << 32)) return fc_fail(ctx, FC_E_ARG, "more than 2^32 records on one device");
    const int parity = (int)((a.barrier_epoch - 1) & 1ull);
    const fc_jrec* half = (const fc_jrec*)a.recs.p + (size_t)parity * (size_t)a.p2p_world * (size_t)a.p2p_slice_cap;
    if (!sort_mode) {
      RecSrc src{half, (const unsigned long long*)a.counters.p + fc::FC_CNT_SLICE + 8 * parity, (unsigned long long)a.p2p_slice_cap,
                 a.p2p_world, (unsigned long long*)a.counters.p};
      // the grid covers the largest slice layout that holds `total` records: thread i serves record i of the concatenation
      const int64_t r = finalize_fused(ctx, total, st, src);
      if (r != -100) return r;
    }
    // sort-based path: it wants one contiguous run
    FC_CUDA(ctx, a.p2p_compact.reserve((size_t)total * sizeof(fc_jrec), st, false, 0));
    int64_t at = 0;
    for (int r = 0; r < a.p2p_world; ++r) {
      if (cnt[r])
        FC_CUDA(ctx, cudaMemcpyAsync((fc_jrec*)a.p2p_compact.p + at, half + (size_t)r * (size_t)a.p2p_slice_cap, (size_t)cnt[r] * sizeof(fc_jrec),
                                     cudaMemcpyDeviceToDevice, st));
      at += (int64_t)cnt[r];
    }
    rbuf = (fc_jrec*)a.p2p_compact.p;
    a.unordered = true;
  } else {
    if (a.n_recs >= (1ll << 32)) return fc_fail(ctx, FC_E_ARG, "more than 2^32 records on one device");
    if (!sort_mode) {
      // no host round trip before the kernels: they are sized by the upper bound and read the exact count on the device
      RecSrc src{(const fc_jrec*)a.recs.p, (const unsigned long long*)a.counters.p, (unsigned long long)a.n_recs, 1, nullptr};
      const int64_t r = finalize_fused(ctx, a.n_recs, st, src);
      if (r != -100) return r;
    }
    int rc = sync_n_recs(ctx, st);
    if (rc) return rc;
    unsigned long long no_name = 0;
    FC_CUDA(ctx, cudaMemcpyAsync(&no_name, (unsigned long long*)a.counters.p + 6, sizeof(no_name), cudaMemcpyDeviceToHost, st));
    FC_CUDA(ctx, cudaStreamSynchronize(st));
    if (no_name)
      return fc_fail(ctx, FC_E_ARG, "%llu records came without a read-name hash and without fragment fields in their descriptors", no_name);
  }
  int rc = FC_OK;
  const int64_t n = a.n_recs;
  if (n == 0) {
    a.n_junc = 0;
    return 0;
  }
  // scratch layout
  FC_CUDA(ctx, a.scratch[0].reserve((size_t)n * 8, st, false, 0));  // u64 A
  FC_CUDA(ctx, a.scratch[1].reserve((size_t)n * 8, st, false, 0));  // u64 B
  FC_CUDA(ctx, a.scratch[2].reserve((size_t)n * 4, st, false, 0));  // u32 A
  FC_CUDA(ctx, a.scratch[3].reserve((size_t)n * 4, st, false, 0));  // u32 B
  FC_CUDA(ctx, a.scratch[4].reserve((size_t)n * sizeof(fc_jrec), st, false, 0));  // sorted records
  FC_CUDA(ctx, a.scratch[5].reserve((size_t)n * 4, st, false, 0));  // head flags
  FC_CUDA(ctx, a.scratch[6].reserve((size_t)n * 4, st, false, 0));  // inclusive segment ids
  uint64_t* kA = (uint64_t*)a.scratch[0].p;
  uint64_t* kB = (uint64_t*)a.scratch[1].p;
  uint32_t* vA = (uint32_t*)a.scratch[2].p;
  uint32_t* vB = (uint32_t*)a.scratch[3].p;
  fc_jrec* sorted = (fc_jrec*)a.scratch[4].p;
  uint32_t* head = (uint32_t*)a.scratch[5].p;
  uint32_t* seg_incl = (uint32_t*)a.scratch[6].p;
  unsigned long long* counters = (unsigned long long*)a.counters.p;

  if (a.unordered) {
    // records arrived in arbitrary order (peer-to-peer emit): restore stream order first, the stable sort below and the
    // sequential float sums rely on it
    idx_key_kernel<<<nblk(n, 256), 256, 0, st>>>(n, (const fc_jrec*)rbuf, kA, vA);
    FC_LAUNCH_CHECK(ctx);
    rc = sort_pairs_u64_u32(ctx, n, kA, kB, vA, vB, 0, 64, st);
    if (rc) return rc;
    gather_kernel<<<nblk(n, 256), 256, 0, st>>>(n, (const fc_jrec*)rbuf, vB, sorted);
    FC_LAUNCH_CHECK(ctx);
    FC_CUDA(ctx, cudaMemcpyAsync(rbuf, sorted, (size_t)n * sizeof(fc_jrec), cudaMemcpyDeviceToDevice, st));
    a.unordered = false;
  }
  uint64_t seed = 0x9E3779B97F4A7C15ULL;
  bool ok = false;
  uint32_t nj32 = 0;
  // sort only as many hash bits as make a collision between two different keys unlikely (~2^-7); the run check
  // below detects one, and the retries use all 64 bits
  int lg = 1;
  while ((1ll << lg) < n) lg++;
  int bits = 2 * lg + 6;
  if (bits > 64) bits = 64;
  for (int attempt = 0; attempt < 4 && !ok; ++attempt, seed = fc_mix64(seed + attempt), bits = 64) {
    const uint64_t hmask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
    key_hash_kernel<<<nblk(n, 256), 256, 0, st>>>(n, (const fc_jrec*)rbuf, seed, kA, vA);
    FC_LAUNCH_CHECK(ctx);
    rc = sort_pairs_u64_u32(ctx, n, kA, kB, vA, vB, 0, bits, st);
    if (rc) return rc;
    gather_kernel<<<nblk(n, 256), 256, 0, st>>>(n, (const fc_jrec*)rbuf, vB, sorted);
    FC_LAUNCH_CHECK(ctx);
    FC_CUDA(ctx, cudaMemsetAsync(counters + 1, 0, sizeof(unsigned long long), st));
    heads_kernel<<<nblk(n, 256), 256, 0, st>>>(n, sorted, kB, hmask, head, counters + 1);
    FC_LAUNCH_CHECK(ctx);
    // segment ids; the collision count and the number of junctions come back with ONE synchronisation
    rc = scan_u32(ctx, n, head, seg_incl, true, st);
    if (rc) return rc;
    unsigned long long coll = 0;
    FC_CUDA(ctx, cudaMemcpyAsync(&coll, counters + 1, sizeof(coll), cudaMemcpyDeviceToHost, st));
    FC_CUDA(ctx, cudaMemcpyAsync(&nj32, seg_incl + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    FC_CUDA(ctx, cudaStreamSynchronize(st));
    ok = coll == 0;
  }
  if (!ok) return fc_fail(ctx, FC_E_COLLISION, "junction key hash collisions survived 4 seeds");
  const int64_t nj = nj32;

  FC_CUDA(ctx, a.scratch[7].reserve((size_t)nj * sizeof(JAcc), st, false, 0));
  JAcc* acc = (JAcc*)a.scratch[7].p;
  acc_init_kernel<<<nblk(nj, 256), 256, 0, st>>>(nj, acc);
  FC_LAUNCH_CHECK(ctx);
  reduce_kernel<<<nblk(n, 256), 256, 0, st>>>(n, sorted, seg_incl, head, acc);
  FC_LAUNCH_CHECK(ctx);

  // distinct read sequences / fragment names per junction: two exact hash sets of (value, junction) pairs
  {
    unsigned long long cap = 1024;
    while (cap < 2ull * (unsigned long long)n) cap <<= 1;
    for (int k = 0; k < 2; ++k) {
      FC_CUDA(ctx, a.htab[k].reserve((size_t)cap * 16, st, false, 0));
      FC_CUDA(ctx, cudaMemsetAsync(a.htab[k].p, 0, (size_t)cap * 16, st));
    }
    distinct_hash_kernel<<<nblk(n, 256), 256, 0, st>>>(n, sorted, seg_incl, (U128*)a.htab[0].p, (U128*)a.htab[1].p, cap - 1, acc);
    FC_LAUNCH_CHECK(ctx);
  }

  FC_CUDA(ctx, a.junctions.reserve((size_t)nj * sizeof(fc_junction), st, false, 0));
  // finish into scratch, then order by first_idx
  FC_CUDA(ctx, a.scratch[5].reserve((size_t)nj * sizeof(fc_junction), st, false, 0));
  fc_junction* tmpj = (fc_junction*)a.scratch[5].p;
  finish_kernel<<<nblk(nj, 128), 128, 0, st>>>(nj, n, acc, sorted, tmpj, kA, vA);
  FC_LAUNCH_CHECK(ctx);
  int order_bits = 64;
  if (a.max_idx != ~0ull) {
    order_bits = 1;
    while (order_bits < 64 && (a.max_idx >> order_bits)) order_bits++;
  }
  rc = sort_pairs_u64_u32(ctx, nj, kA, kB, vA, vB, 0, order_bits, st);
  if (rc) return rc;
  gather_junctions_kernel<<<nblk(nj, 256), 256, 0, st>>>(nj, tmpj, vB, (fc_junction*)a.junctions.p);
  FC_LAUNCH_CHECK(ctx);
  a.n_junc = nj;
  return nj;
}

extern "C" int fc_agg_set_idx_range(fc_ctx* ctx, uint64_t lo, uint64_t hi) {
  if (!ctx || hi <= lo) return FC_E_ARG;
  ctx->agg.idx_lo = lo;
  ctx->agg.max_idx = hi;
  ctx->agg.range_declared = true;
  return FC_OK;
}

extern "C" int fc_agg_set_timing(fc_ctx* ctx, int32_t on) {
  if (!ctx) return FC_E_ARG;
  ctx->agg.timing = on != 0;
  for (float& v : ctx->agg.stage_us) v = 0.f;
  return FC_OK;
}

extern "C" int fc_agg_get_timing(fc_ctx* ctx, float* out_us /* 6 */) {
  if (!ctx || !out_us) return FC_E_ARG;
  for (int k = 0; k < 6; ++k) out_us[k] = ctx->agg.stage_us[k];
  return FC_OK;
}

extern "C" int fc_agg_fetch(fc_ctx* ctx, int64_t n, fc_junction* h_out) {
  if (!ctx || !h_out) return FC_E_ARG;
  if (ctx->agg.n_junc < 0) return fc_fail(ctx, FC_E_STATE, "fc_agg_fetch before fc_agg_finalize");
  if (n > ctx->agg.n_junc) n = ctx->agg.n_junc;
  if (n <= 0) return FC_OK;
  FC_CUDA(ctx, cudaDeviceSynchronize());
  FC_CUDA(ctx, cudaMemcpy(h_out, ctx->agg.junctions.p, (size_t)n * sizeof(fc_junction), cudaMemcpyDeviceToHost));
  return FC_OK;
}

extern "C" const fc_junction* fc_agg_junctions(fc_ctx* ctx) {
  return ctx && ctx->agg.n_junc >= 0 ? (const fc_junction*)ctx->agg.junctions.p : nullptr;
}

// ======================================================================================================================
// Fused emit + exchange over peer memory (multi-GPU, one node).  Every rank exports its record buffer and its counter
// block (CUDA IPC between processes; plain device pointers between contexts of one process); the emit kernel of every rank
// hashes the junction key of each accepted pair to its owner rank and stores the 48-byte record straight into the slice
// that the owner's buffer reserves for this source (emit_core.cuh) -- slots come from counters in the SOURCE's memory, so
// no atomic crosses NVLink.  One stream-ordered barrier per step publishes the slice counts to the owners and orders the
// record stores before the owners' fc_agg_finalize; the buffers have two halves used by alternating steps, so no barrier
// is needed before the next step's stores.  Arrival order is arbitrary; the sort-free aggregation does not depend on it and
// the sort-based fallback re-orders by idx.
// ======================================================================================================================
using fc::P2PView;
constexpr int P2P_THREADS = 512;

__global__ void __launch_bounds__(P2P_THREADS) emit_p2p_kernel(int64_t n, const fc_hit* __restrict__ hits,
                                                               const uint8_t* __restrict__ mask,
                                                               const int32_t* __restrict__ chrom, const uint8_t* __restrict__ flags,
                                                               fc::EmitArgs e, P2PView pv, unsigned long long* __restrict__ overflow) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  fc_hit h = {0, 0, 0u, 0u};
  bool accept = false;
  uint32_t c = 0, fl = 0;
  if (i < n) {
    h = hits[i];
    accept = (h.w2 & 0xFFFFu) != 0u && (!mask || mask[i]);
    if (accept) {
      c = (uint32_t)chrom[i];
      fl = flags[i];
    }
  }
  fc_jrec r;
  if (accept) r = fc::make_record(i, h.start, h.end, h.w2, h.w3, c, fl, e);
  fc::emit_p2p_block<P2P_THREADS>(accept, r, pv, overflow);
}

static int p2p_reserve(fc_ctx* ctx, int64_t capacity_records) {
  fc_agg& a = ctx->agg;
  cudaStream_t st = ctx->own_stream;
  int rc = ensure_counters(ctx, st);
  if (rc) return rc;
  // two halves (alternating steps) of `capacity` records each
  FC_CUDA(ctx, a.recs.reserve(2 * (size_t)capacity_records * sizeof(fc_jrec), st, false, 0));
  FC_CUDA(ctx, cudaStreamSynchronize(st));
  a.p2p_capacity = capacity_records;
  return FC_OK;
}

extern "C" int fc_p2p_export(fc_ctx* ctx, int64_t capacity_records, uint8_t* h_handles /* 128 bytes */) {
  if (!ctx || capacity_records <= 0 || !h_handles) return FC_E_ARG;
  fc_agg& a = ctx->agg;
  int rc = p2p_reserve(ctx, capacity_records);
  if (rc) return rc;
  cudaIpcMemHandle_t h0, h1;
  FC_CUDA(ctx, cudaIpcGetMemHandle(&h0, a.recs.p));
  FC_CUDA(ctx, cudaIpcGetMemHandle(&h1, a.counters.p));
  memcpy(h_handles, &h0, 64);
  memcpy(h_handles + 64, &h1, 64);
  return FC_OK;
}

// the same for contexts of ONE process on one device: plain device pointers instead of IPC handles
extern "C" int fc_p2p_export_local(fc_ctx* ctx, int64_t capacity_records, void** out_recs, void** out_counters) {
  if (!ctx || capacity_records <= 0 || !out_recs || !out_counters) return FC_E_ARG;
  int rc = p2p_reserve(ctx, capacity_records);
  if (rc) return rc;
  *out_recs = ctx->agg.recs.p;
  *out_counters = ctx->agg.counters.p;
  return FC_OK;
}

static int p2p_finish_connect(fc_ctx* ctx, int32_t world, int32_t rank, const int64_t* h_capacities) {
  fc_agg& a = ctx->agg;
  int64_t cap = a.p2p_capacity;
  for (int r = 0; r < world; ++r)
    if (h_capacities[r] < cap) cap = h_capacities[r];
  a.p2p_world = world;
  a.p2p_rank = rank;
  a.p2p_slice_cap = cap / world;
  if (a.p2p_slice_cap < 1) return fc_fail(ctx, FC_E_ARG, "peer-to-peer capacity %lld too small for %d ranks", (long long)cap, world);
  a.p2p_enabled = true;
  a.p2p_open = false;
  a.barrier_epoch = 0;
  return FC_OK;
}

extern "C" int fc_p2p_connect(fc_ctx* ctx, int32_t world, int32_t rank, const uint8_t* h_all_handles /* world x 128 */,
                              const int64_t* h_capacities /* world */) {
  if (!ctx || world < 1 || world > 8 || rank < 0 || rank >= world || !h_all_handles || !h_capacities) return FC_E_ARG;
  fc_agg& a = ctx->agg;
  if (a.p2p_capacity <= 0) return fc_fail(ctx, FC_E_STATE, "fc_p2p_connect before fc_p2p_export");
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      a.p2p_recs[r] = a.recs.p;
      a.p2p_cnt[r] = a.counters.p;
      continue;
    }
    cudaIpcMemHandle_t h0, h1;
    memcpy(&h0, h_all_handles + (size_t)r * 128, 64);
    memcpy(&h1, h_all_handles + (size_t)r * 128 + 64, 64);
    FC_CUDA(ctx, cudaIpcOpenMemHandle(&a.p2p_recs[r], h0, cudaIpcMemLazyEnablePeerAccess));
    FC_CUDA(ctx, cudaIpcOpenMemHandle(&a.p2p_cnt[r], h1, cudaIpcMemLazyEnablePeerAccess));
  }
  a.p2p_ipc = true;
  a.p2p_local = false;
  return p2p_finish_connect(ctx, world, rank, h_capacities);
}

extern "C" int fc_p2p_connect_local(fc_ctx* ctx, int32_t world, int32_t rank, void* const* recs, void* const* counters,
                                    const int64_t* h_capacities) {
  if (!ctx || world < 1 || world > 8 || rank < 0 || rank >= world || !recs || !counters || !h_capacities) return FC_E_ARG;
  fc_agg& a = ctx->agg;
  if (a.p2p_capacity <= 0) return fc_fail(ctx, FC_E_STATE, "fc_p2p_connect_local before fc_p2p_export_local");
  for (int r = 0; r < world; ++r) {
    a.p2p_recs[r] = r == rank ? a.recs.p : recs[r];
    a.p2p_cnt[r] = r == rank ? a.counters.p : counters[r];
  }
  a.p2p_ipc = false;
  a.p2p_local = true;
  return p2p_finish_connect(ctx, world, rank, h_capacities);
}

extern "C" int fc_p2p_set_timeout(fc_ctx* ctx, double seconds) {
  if (!ctx || !(seconds > 0.0)) return FC_E_ARG;
  ctx->agg.p2p_timeout_s = seconds;
  return FC_OK;
}

static void p2p_view(const fc_agg& a, P2PView* pv) {
  for (int r = 0; r < 8; ++r) {
    pv->recs[r] = r < a.p2p_world ? (fc_jrec*)a.p2p_recs[r] : nullptr;
    pv->cnt[r] = r < a.p2p_world ? (unsigned long long*)a.p2p_cnt[r] : nullptr;
  }
  pv->slice_cap = (unsigned long long)a.p2p_slice_cap;
  pv->world = a.p2p_world;
  pv->rank = a.p2p_rank;
  pv->parity = (int)(a.barrier_epoch & 1ull);
}

// the peer view of the context and the bookkeeping of a peer emit, shared with the scan kernel that emits on the way
int fc_agg_p2p_begin(fc_ctx* ctx, fc::P2PView* pv, unsigned long long** overflow) {
  fc_agg& a = ctx->agg;
  if (!a.p2p_enabled) return fc_fail(ctx, FC_E_STATE, "peer emit before fc_p2p_connect");
  p2p_view(a, pv);
  *overflow = (unsigned long long*)a.counters.p + 4;
  return FC_OK;
}

void fc_agg_p2p_end(fc_ctx* ctx, uint64_t idx_base, int64_t n) {
  fc_agg& a = ctx->agg;
  a.p2p_open = true;  // until the barrier that ends the step
  a.n_exact = false;
  a.unordered = true;
  if (!a.range_declared) a.max_idx = ~0ull;
  (void)idx_base;
  (void)n;
  a.n_junc = -1;
}

extern "C" int fc_agg_emit_p2p(fc_ctx* ctx, int64_t n, const fc_hit* d_hits, const int32_t* d_chrom, const uint8_t* d_flags,
                               const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b,
                               const uint64_t* d_read_hash, const uint64_t* d_qname_hash, const uint8_t* d_mask,
                               uint64_t idx_base, void* stream) {
  if (!ctx || n < 0) return FC_E_ARG;
  P2PView pv;
  unsigned long long* overflow = nullptr;
  int rc = fc_agg_p2p_begin(ctx, &pv, &overflow);
  if (rc) return rc;
  if (n == 0) {
    fc_agg_p2p_end(ctx, idx_base, 0);  // an empty shard still owns keys: its peers' records must be reduced
    return FC_OK;
  }
  fc::EmitArgs e{d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, nullptr, nullptr, nullptr};
  emit_p2p_kernel<<<nblk(n, P2P_THREADS), P2P_THREADS, 0, (cudaStream_t)stream>>>(n, d_hits, d_mask, d_chrom, d_flags, e, pv, overflow);
  FC_LAUNCH_CHECK(ctx);
  fc_agg_p2p_end(ctx, idx_base, n);
  return FC_OK;
}

// ---- stream-ordered barrier over peer memory ------------------------------------------------------------------
// Ends a step: every rank (1) publishes how many records it has put into its slice of every owner, (2) adds 1 to the
// arrival word of every rank (its own included) and (3) waits until its own word has seen `world` arrivals per barrier so
// far -- then every peer's records and counts have landed here.  One tiny kernel: a few microseconds over NVLink instead
// of a collective launch.  The wait is bounded (a peer that died must not hang the GPU): after `timeout_cycles` it gives
// up and raises the flag that makes the next fc_agg_finalize fail with FC_E_STATE -- the step's results are then invalid.
// wait == 0 (contexts of one process on one device, launched one after the other): publish and arrive only.
__global__ void p2p_barrier_kernel(P2PView pv, unsigned long long target, long long timeout_cycles, int wait) {
  if ((int)threadIdx.x < pv.world) {
    const unsigned long long sent = pv.cnt[pv.rank][fc::FC_CNT_SRC + fc::FC_CNT_SRC_STRIDE * threadIdx.x];
    pv.cnt[threadIdx.x][fc::FC_CNT_SLICE + 8 * pv.parity + pv.rank] = sent;
    pv.cnt[pv.rank][fc::FC_CNT_SRC + fc::FC_CNT_SRC_STRIDE * threadIdx.x] = 0ull;  // the next step starts from empty slices
  }
  __threadfence_system();  // this rank's record stores and counts are ordered before its arrival
  __syncwarp();
  if ((int)threadIdx.x < pv.world) atomicAdd_system(pv.cnt[threadIdx.x] + FC_BARRIER_WORD, 1ull);
  if (wait && threadIdx.x == 0) {
    volatile unsigned long long* mine = pv.cnt[pv.rank] + FC_BARRIER_WORD;
    const long long t0 = clock64();
    while (*mine < target) {
      if (clock64() - t0 > timeout_cycles) {
        pv.cnt[pv.rank][FC_BARRIER_TIMEOUT_WORD] = 1ull;
        break;
      }
      __nanosleep(100);
    }
    __threadfence_system();
  }
}

extern "C" int fc_p2p_barrier(fc_ctx* ctx, void* stream) {
  if (!ctx) return FC_E_ARG;
  fc_agg& a = ctx->agg;
  if (!a.p2p_enabled) return fc_fail(ctx, FC_E_STATE, "fc_p2p_barrier before fc_p2p_connect");
  P2PView pv;
  p2p_view(a, &pv);
  a.barrier_epoch++;
  a.p2p_open = false;
  const long long cycles = (long long)(a.p2p_timeout_s * 2.0e9);  // clock64 ticks at <= 2 GHz: at least the requested time
  p2p_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pv, a.barrier_epoch * (unsigned long long)a.p2p_world, cycles,
                                                         a.p2p_local ? 0 : 1);
  FC_LAUNCH_CHECK(ctx);
  return FC_OK;
}

void fc_agg_release(fc_ctx* ctx) {
  fc_agg& a = ctx->agg;
  if (a.p2p_ipc)
    for (int r = 0; r < a.p2p_world; ++r)
      if (r != a.p2p_rank) {
        if (a.p2p_recs[r]) cudaIpcCloseMemHandle(a.p2p_recs[r]);
        if (a.p2p_cnt[r]) cudaIpcCloseMemHandle(a.p2p_cnt[r]);
      }
  a.p2p_enabled = false;
  a.p2p_compact.release();
  a.recs.release();
  a.junctions.release();
  for (auto& s : a.scratch) s.release();
  a.cub_tmp.release();
  a.counters.release();
  for (auto& h : a.htab) h.release();
  a.f_keys.release();
  a.f_sets.release();
  a.f_acc.release();
  a.f_part.release();
  a.f_pcur.release();
  if (a.h_pinned) cudaFreeHost(a.h_pinned);
  a.h_pinned = nullptr;
  if (a.side) {
    cudaStreamDestroy(a.side);
    cudaEventDestroy(a.ev_side);
    cudaEventDestroy(a.ev_main);
    a.side = nullptr;
  }
}
