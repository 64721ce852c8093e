// scan_core.cuh -- per-anchor-pair breakpoint scan (device code; also compiles for the host so that the CPU-only
// development container can check it bit-for-bit against the oracle -- tests/tools/scan_host_harness.cpp --
// the shipped library only ever runs it inside CUDA kernels).
//
// Replaces JunctionSpan.find_breakpoints + Splice.score (/root/reference/find_circ.py:766-806, 854-974).
//
// Reference algorithm, for l internal read bases R[0..l) and two (l+2)-base genome windows A (donor side,
// starting at A.pos+eff) and B (acceptor side, ending at B.aend-eff):
//     for x in 0..l:  dist(x) = #{i<x : A[i]!=R[i]} + #{i>=x : B[i+2]!=R[i]}            (:906-908, O(l) bytes each -> O(l^2))
//                     keep if dist<=maxdist and (A[x],A[x+1],B[x],B[x+1]) is GT/AG ('+') or CT/AC ('-')   (:915-954)
//     rank by 20*canonical - 10*dist - ov (+100*strand match), stable, ties counted         (:792-799, 961-974)
//
// B200 formulation -- one thread per pair, everything in registers, 1 bit per base and plane:
//   * the genome is held as bit planes (lo / hi bit of the 2-bit code, plus an N plane).  For the scan the planes
//     are re-cut into overlapping 32*T-byte TILES starting every 32 bases (every window lies inside one tile), the
//     lo and hi plane of a tile sit in the same sector(s) and 4 spare bits flag "tile contains N": one 256-bit
//     load per sector fetches a whole window including its N summary -- two gathers per pair when T = 1;
//   * mismatch flags of 32 bases cost two XORs and an OR; dist(x) = popc(mA below x) + popc(mB at/above x);
//   * GT/AG and CT/AC positions are found for all split positions at once with plane logic, and a word is only
//     searched when the mismatches before it (donor side) and after it (acceptor side) both stay <= maxdist,
//     so chance signals far from the true breakpoint are never scored;
//   * windows or reads containing N take the same bit-parallel code with the N planes loaded from the master store
//     (bytes are compared as the reference does: N equals N and differs from every base);
//   * --non-canonical and windows longer than 256 bases take an exact per-base loop.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FC_HD __host__ __device__ __forceinline__
#else
#define FC_HD inline
#endif

namespace fc {

struct GenomeView {
  // master store: 1 bit per base and plane, base g at word g>>5, bit g&31
  const uint32_t* plo;
  const uint32_t* phi;
  const uint32_t* pn;  // 1 = not ACGT (reads as 'N'); the padding between chromosomes is N
  const int64_t* chrom_off;  // global index of base 0 of each chromosome
  const int64_t* chrom_size;
  int32_t n_chrom;
  int32_t pad;  // bases of N padding on both sides of every chromosome
  // tile store (may be absent: tile_T == 0)
  const uint32_t* tiles;   // tile t: 4T words lo plane (top 4 bits of the last word = flags), then 4T words hi plane
  int32_t tile_T;          // sectors per tile (1, 2 or 4)
  int32_t tile_W;          // largest window (bases) guaranteed to fit in one tile
};

struct ScanCfg {
  int32_t margin, maxdist, noncanonical, strandpref;
};

struct HitOut {
  int32_t start, end;
  uint32_t w2, w3;
};

constexpr uint32_t SIG_GTAG = 2u | (3u << 3) | (0u << 6) | (2u << 9);
constexpr uint32_t W3_RANGE = 1u << 30;
constexpr uint32_t W3_SLOW = 1u << 31;
constexpr uint32_t TILE_FLAG_N = 1u << 31;

FC_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
FC_HD int ffs32(uint32_t x) {  // index of lowest set bit, x != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
FC_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {  // (hi:lo >> s), s in [0,32)
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
FC_HD int imin(int a, int b) { return a < b ? a : b; }
FC_HD int imax(int a, int b) { return a > b ? a : b; }
FC_HD uint32_t low_mask(int k) {  // k in [0,32]: the k lowest bits
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(0xFFFFFFFFu, 0u, 32 - k);
#else
  return k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
#endif
}
FC_HD uint32_t ldg32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
FC_HD uint64_t umul64hi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// one 32-byte sector (8 words), p 32-byte aligned
FC_HD void ldg256(const uint32_t* p, uint32_t (&v)[8]) {
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
#else
  for (int k = 0; k < 8; ++k) v[k] = p[k];
#endif
}

// ---------------------------------------------------------------- running best / tie bookkeeping
struct Best {
  int score;     // best score so far
  int n_ties;    // hits with that score
  int x;         // split position of the first best hit
  uint32_t info; // strand | sig<<1 | dist<<16 | ov<<24 of the first best hit
  FC_HD void init() {
    score = -100000;
    n_ties = 0;
    x = -1;
    info = 0;
  }
  // hits must be offered in the reference's list order (ascending x, '+' before '-')
  FC_HD void offer(int s, int xx, uint32_t strand, uint32_t sig, int dist, int ov) {
    if (s > score) {
      score = s;
      n_ties = 1;
      x = xx;
      info = strand | (sig << 1) | ((uint32_t)dist << 16) | ((uint32_t)ov << 24);
    } else if (s == score) {
      n_ties++;
    }
  }
};

FC_HD int anchor_overlap(int x, int l, int margin) {  // find_circ.py:917-922 (second test overrides the first)
  int ov = 0;
  if (margin) {
    if (x < margin) ov = margin - x;
    if (l - x < margin) ov = margin - (l - x);
  }
  return ov;
}

FC_HD void finish(const Best& b, int a_start, int b_end, int l, bool backsplice, uint32_t extra, HitOut& out) {
  if (b.n_ties == 0) {
    out.start = 0;
    out.end = 0;
    out.w2 = 0;
    out.w3 = extra;
    return;
  }
  // find_circ.py:929-945
  int s = b_end - l + b.x, e = a_start + b.x + 1;
  int lo = s < e ? s : e, hi = s < e ? e : s;
  if (backsplice)
    hi -= 1;
  else
    lo -= 1;
  out.start = lo;
  out.end = hi;
  uint32_t nh = b.n_ties > 65535 ? 65535u : (uint32_t)b.n_ties;
  out.w2 = nh | (b.info & 0xFFFF0000u);
  out.w3 = (b.info & 0x1FFFu) | ((uint32_t)((b.score + 512) & 1023) << 13) | extra;
}

// ---------------------------------------------------------------- exact per-base path
FC_HD int gcode(const GenomeView& g, int64_t gp) {  // 0..3, 4 = N
  const int64_t w = gp >> 5;
  const uint32_t sh = (uint32_t)(gp & 31);
  if ((ldg32(g.pn + w) >> sh) & 1u) return 4;
  return (int)(((ldg32(g.plo + w) >> sh) & 1u) | (((ldg32(g.phi + w) >> sh) & 1u) << 1));
}

// Read planes of a batch: word w of pair i of the lo / hi plane at [i * pair_stride + w * stride], of the N plane at
// [i * rn_pair_stride + w * stride]; base j in word j>>5, bit j&31.
//   word-major (fc_pairs):  pair_stride = rn_pair_stride = 1, stride = plane stride (>= n)
//   row-major  (fc_batch):  rows [lo words | hi words], rhi = rlo + NW, pair_stride = 2 NW, rn_pair_stride = NW, stride = 1
struct ReadView {
  const uint32_t* rlo;
  const uint32_t* rhi;
  const uint32_t* rn;
  int64_t n;       // pairs
  int32_t n_words; // words per pair and plane
  int64_t stride;  // distance (in words) between word w and word w+1 of one pair
  int64_t pair_stride = 1;
  int64_t rn_pair_stride = 1;
};

FC_HD int rcode(const ReadView& rv, int64_t i, int j, bool has_n) {
  const int64_t w = (int64_t)(j >> 5) * rv.stride;
  const int64_t a = w + i * rv.pair_stride;
  const uint32_t sh = (uint32_t)(j & 31);
  if (has_n && ((ldg32(rv.rn + w + i * rv.rn_pair_stride) >> sh) & 1u)) return 4;
  return (int)(((ldg32(rv.rlo + a) >> sh) & 1u) | (((ldg32(rv.rhi + a) >> sh) & 1u) << 1));
}
FC_HD uint32_t comp_code(uint32_t c) { return c == 4 ? 4u : 3u - c; }

// Emit: functor called for every hit in reference list order; used by the tie enumerator.
struct NoEmit {
  FC_HD void operator()(int, int, uint32_t, uint32_t, int, int) const {}
};

template <class Emit>
FC_HD void scan_per_base(const GenomeView& g, const ScanCfg& cfg, int64_t ga, int64_t gb, int l, bool minus_span,
                         const ReadView& rv, int64_t i, bool has_n, Best& best, Emit& emit) {
  // dist(0): everything explained by the acceptor side
  int dist = 0;
  for (int j = 0; j < l; ++j) dist += gcode(g, gb + j + 2) != rcode(rv, i, j, has_n);
  for (int x = 0; x <= l; ++x) {
    if (dist <= cfg.maxdist) {
      uint32_t a0 = gcode(g, ga + x), a1 = gcode(g, ga + x + 1), b0 = gcode(g, gb + x), b1 = gcode(g, gb + x + 1);
      uint32_t sig = a0 | (a1 << 3) | (b0 << 6) | (b1 << 9);
      uint32_t rc = comp_code(b1) | (comp_code(b0) << 3) | (comp_code(a1) << 6) | (comp_code(a0) << 9);
      int ov = anchor_overlap(x, l, cfg.margin);
      int base = -10 * dist - ov;
      bool is_gtag = sig == SIG_GTAG, is_ctac = rc == SIG_GTAG;
      if (cfg.noncanonical) {
        // noncanonical == 2: the v1.2 ranking (edits, then anchor overlap; README.md:301-312) without the canonical bonus
        const int bonus = cfg.noncanonical == 2 ? 0 : 20;
        int sp = base + (is_gtag ? bonus : 0) + ((cfg.strandpref && !minus_span) ? 100 : 0);
        int sm = base + (is_ctac ? bonus : 0) + ((cfg.strandpref && minus_span) ? 100 : 0);
        best.offer(sp, x, 0u, sig, dist, ov);
        emit(sp, x, 0u, sig, dist, ov);
        best.offer(sm, x, 1u, rc, dist, ov);
        emit(sm, x, 1u, rc, dist, ov);
      } else if (is_gtag) {
        int sp = base + 20 + ((cfg.strandpref && !minus_span) ? 100 : 0);
        best.offer(sp, x, 0u, SIG_GTAG, dist, ov);
        emit(sp, x, 0u, SIG_GTAG, dist, ov);
      } else if (is_ctac) {
        int sm = base + 20 + ((cfg.strandpref && minus_span) ? 100 : 0);
        best.offer(sm, x, 1u, SIG_GTAG, dist, ov);
        emit(sm, x, 1u, SIG_GTAG, dist, ov);
      }
    }
    if (x < l) {
      int r = rcode(rv, i, x, has_n);
      dist += (gcode(g, ga + x) != r) - (gcode(g, gb + x + 2) != r);
    }
  }
}

// ---------------------------------------------------------------- window loaders
// A window is NP words per plane (32 bases per word) + one zero guard word.
template <int NP>
struct Window {
  uint32_t lo[NP + 1], hi[NP + 1];
};

// from the master planes: any alignment, NP+1 scalar loads per plane (rare path)
template <int NP>
FC_HD void load_plane(const uint32_t* plane, int64_t gp, uint32_t (&W)[NP + 1]) {
  const uint32_t* base = plane + (gp >> 5);
  const uint32_t bo = (uint32_t)(gp & 31);
  uint32_t prev = ldg32(base);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    uint32_t next = ldg32(base + j + 1);
    W[j] = funnel_r(prev, next, bo);
    prev = next;
  }
  W[NP] = 0;
}

// from the tile store: one 256-bit load per sector.  Tiles start every 32 bases (tile t covers bases [32t, 32t+P)),
// so the tile index is a shift and the window starts inside the tile's first word: no division, no word select.
// Returns the tile flags (TILE_FLAG_N).
template <int NP, int T>
FC_HD uint32_t load_tile_window(const GenomeView& g, int64_t gp, Window<NP>& w) {
  constexpr int PW = 4 * T;  // words per plane in a tile
  static_assert(NP <= PW, "window does not fit the tile");
  const uint32_t bo = (uint32_t)(gp & 31);
  const uint32_t* base = g.tiles + (gp >> 5) * (int64_t)(8 * T);
  uint32_t s_lo[PW + 1], s_hi[PW + 1];
  if (T == 1) {
    uint32_t v[8];
    ldg256(base, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s_lo[k] = v[k];
      s_hi[k] = v[4 + k];
    }
  } else {
#pragma unroll
    for (int q = 0; q < T / 2; ++q) {
      uint32_t v[8];
      ldg256(base + 8 * q, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s_lo[8 * q + k] = v[k];
      ldg256(base + PW + 8 * q, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s_hi[8 * q + k] = v[k];
    }
  }
  const uint32_t flags = s_lo[PW - 1] & 0xF0000000u;
  s_lo[PW - 1] &= 0x0FFFFFFFu;
  s_lo[PW] = 0;
  s_hi[PW] = 0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    w.lo[j] = funnel_r(s_lo[j], s_lo[j + 1], bo);
    w.hi[j] = funnel_r(s_hi[j], s_hi[j + 1], bo);
  }
  w.lo[NP] = 0;
  w.hi[NP] = 0;
  return flags;
}

// ---------------------------------------------------------------- bit-parallel scan on loaded windows
// NP words of 32 split positions; covers l + 2 <= 32*NP.  WITH_N is chosen per WARP (any lane touching an N), so an
// N-free warp runs the short variant and a warp with N runs the full one for all its lanes -- never both.  In the
// full variant nA/nB are the N planes of the windows (all zero for the lanes without N).
template <int NP, bool WITH_N>
FC_HD void scan_planes(const ScanCfg& cfg, const Window<NP>& A, const Window<NP>& B, const uint32_t (&nA)[NP + 1],
                       const uint32_t (&nB)[NP + 1], int l, bool minus_span, const uint32_t (&rlo)[NP],
                       const uint32_t (&rhi)[NP], const uint32_t (&rnn)[NP], Best& best) {
  // splice signal at split position x (bit x&31 of word x>>5), codes A=00 C=01 G=10 T=11 (hi,lo):
  //   GT..AG: A[x]=G A[x+1]=T B[x]=A B[x+1]=G ;  CT..AC: A[x]=C A[x+1]=T B[x]=A B[x+1]=C     (find_circ.py:924-954)
  //   <=> A[x+1]==T, B[x]==A, A[x]==B[x+1], A[x] in {C,G};  '-' strand iff A[x]==C (lo bit set); never contains N
  uint32_t sig[NP];
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const uint32_t a1lo = funnel_r(A.lo[k], A.lo[k + 1], 1), a1hi = funnel_r(A.hi[k], A.hi[k + 1], 1);
    const uint32_t b1lo = funnel_r(B.lo[k], B.lo[k + 1], 1), b1hi = funnel_r(B.hi[k], B.hi[k + 1], 1);
    uint32_t v = (a1hi & a1lo) & ~(B.hi[k] | B.lo[k]);
    v &= (A.hi[k] ^ A.lo[k]) & ~(A.hi[k] ^ b1hi) & ~(A.lo[k] ^ b1lo);
    v &= low_mask(imin(imax(l + 1 - 32 * k, 0), 32));
    if (WITH_N) v &= ~(nA[k] | funnel_r(nA[k], nA[k + 1], 1) | nB[k] | funnel_r(nB[k], nB[k + 1], 1));
    sig[k] = v;
    any |= v;
  }
  if (!any) return;  // no split position carries a canonical signal (most decoy pairs end here)

  // mismatch flags of the read against the donor window (A[i] vs R[i]) and the acceptor window (B[i+2] vs R[i]).
  // N is stored as code 0 on both sides: the base difference is 0 where both are N, the XOR of the N flags decides
  // the rest (N equals N, N differs from every base -- the reference compares bytes, find_circ.py:861-863)
  uint32_t mA[NP], mB[NP];
  int cumA[NP + 1], cumB[NP + 1];
  cumA[0] = 0;
  cumB[0] = 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const uint32_t b2lo = funnel_r(B.lo[k], B.lo[k + 1], 2), b2hi = funnel_r(B.hi[k], B.hi[k + 1], 2);
    uint32_t fa = (A.lo[k] ^ rlo[k]) | (A.hi[k] ^ rhi[k]);
    uint32_t fb = (b2lo ^ rlo[k]) | (b2hi ^ rhi[k]);
    if (WITH_N) {
      fa |= nA[k] ^ rnn[k];
      fb |= funnel_r(nB[k], nB[k + 1], 2) ^ rnn[k];
    }
    const uint32_t vm = low_mask(imin(imax(l - 32 * k, 0), 32));
    mA[k] = fa & vm;
    mB[k] = fb & vm;
    cumA[k + 1] = cumA[k] + popc32(mA[k]);
    cumB[k + 1] = cumB[k] + popc32(mB[k]);
  }
  const int totalB = cumB[NP];

  // candidates: signal positions of the words that can still reach dist <= maxdist (a split inside word k has at
  // least cumA[k] donor-side and totalB - cumB[k+1] acceptor-side mismatches)
  any = 0;
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const bool feasible = cumA[k] <= cfg.maxdist && totalB - cumB[k + 1] <= cfg.maxdist;
    sig[k] = feasible ? sig[k] : 0u;
    any |= sig[k];
  }
  // ONE loop over all words (ascending x), so that the lanes of a warp walk their candidates together
  while (any) {
    int k = 0;
    uint32_t cw = 0, ma = 0, mb = 0, alo = 0;
    int ca = 0, cb = 0;
#pragma unroll
    for (int q = NP - 1; q >= 0; --q) {
      if (sig[q]) {
        k = q;
        cw = sig[q];
        ma = mA[q];
        mb = mB[q];
        alo = A.lo[q];
        ca = cumA[q];
        cb = cumB[q];
      }
    }
    const int bit = ffs32(cw);
    const uint32_t rest = cw & (cw - 1u);
    any = 0;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      if (q == k) sig[q] = rest;
      any |= sig[q];
    }
    const uint32_t below = (1u << bit) - 1u;
    const int dist = ca + popc32(ma & below) + (totalB - cb - popc32(mb & below));
    if (dist <= cfg.maxdist) {
      const int x = 32 * k + bit;
      const uint32_t strand = (alo >> bit) & 1u;  // A[x]==C -> CT..AC -> '-'
      const int ov = anchor_overlap(x, l, cfg.margin);
      int s = 20 - 10 * dist - ov;
      if (cfg.strandpref && ((strand != 0u) == minus_span)) s += 100;
      best.offer(s, x, strand, SIG_GTAG, dist, ov);
    }
  }
}

// ---------------------------------------------------------------- windows by global coordinate, read planes in registers
FC_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
  return __any_sync(__activemask(), p);
#else
  return p;
#endif
}

// NP: 32-base words per window held in registers; T: sectors per tile of the tile store (0: no tile store).
// ga / gb: global coordinates of the first base of the donor / acceptor window (l + 2 <= 32 NP bases each).
template <int NP, int T>
FC_HD void scan_fast(const GenomeView& g, const ScanCfg& cfg, int64_t ga, int64_t gb, int l, bool minus_span, bool read_n,
                     const uint32_t (&rlo)[NP], const uint32_t (&rhi)[NP], const ReadView& rv, int64_t i, Best& best) {
  Window<NP> A, B;
  uint32_t nA[NP + 1], nB[NP + 1], rnn[NP];
  bool with_n = read_n;
  bool tiled = false;
  if constexpr (T > 0) {
    if (l + 2 <= g.tile_W) {
      const uint32_t fl = load_tile_window<NP, T>(g, ga, A) | load_tile_window<NP, T>(g, gb, B);
      with_n = with_n || (fl & TILE_FLAG_N);
      tiled = true;
    }
  }
  if (!tiled) {
    load_plane<NP>(g.plo, ga, A.lo);
    load_plane<NP>(g.phi, ga, A.hi);
    load_plane<NP>(g.plo, gb, B.lo);
    load_plane<NP>(g.phi, gb, B.hi);
    with_n = true;  // no summary without tiles: always consult the N plane
  }
  // one variant per warp (the lanes that reached this point): the N-aware one only if some lane needs it
  if (warp_any(with_n)) {
#pragma unroll
    for (int k = 0; k < NP; ++k)
      rnn[k] = (read_n && k < rv.n_words) ? ldg32(rv.rn + (int64_t)k * rv.stride + i * rv.rn_pair_stride) : 0u;
    if (with_n) {
      load_plane<NP>(g.pn, ga, nA);
      load_plane<NP>(g.pn, gb, nB);
    } else {
#pragma unroll
      for (int k = 0; k <= NP; ++k) nA[k] = nB[k] = 0u;
    }
    scan_planes<NP, true>(cfg, A, B, nA, nB, l, minus_span, rlo, rhi, rnn, best);
  } else {
    scan_planes<NP, false>(cfg, A, B, nA, nB, l, minus_span, rlo, rhi, rnn, best);
  }
}

// ---------------------------------------------------------------- one pair
struct PairArgs {
  int32_t chrom, a_start, b_end, l;
  uint32_t flags;  // FC_PF_*
};

template <int NP, int T, class Emit>
FC_HD void scan_pair(const GenomeView& g, const ScanCfg& cfg, const PairArgs& p, const ReadView& rv, int64_t i,
                     HitOut& out, Emit& emit, bool force_per_base) {
  Best best;
  best.init();
  const bool backsplice = p.flags & 1u, minus_span = p.flags & 2u, read_n = p.flags & 4u;
  const int l = p.l;
  uint32_t extra = 0;
  bool fast = false;
  int64_t ga = 0, gb = 0;
  if (l >= 0 && p.chrom >= 0 && p.chrom < g.n_chrom && l + 2 <= g.pad) {
    const int64_t size = g.chrom_size[p.chrom];
    const int w = l + 2;
    // both windows must touch the chromosome (find_circ.py:194-211 pads the overhang with N; a window entirely
    // outside is undefined there): -w <= a_start <= size  and  0 <= b_end <= size + w   (w <= pad keeps the overhang
    // inside the N padding)
    const bool ok = (uint64_t)((int64_t)p.a_start + w) <= (uint64_t)(size + w) && (uint64_t)(int64_t)p.b_end <= (uint64_t)(size + w);
    if (ok) {
      const int64_t off = g.chrom_off[p.chrom];
      ga = off + p.a_start;
      gb = off + (int64_t)p.b_end - w;
      if (force_per_base || cfg.noncanonical || (w > 32 * NP)) {
        extra |= W3_SLOW;
        scan_per_base(g, cfg, ga, gb, l, minus_span, rv, i, read_n, best, emit);
      } else {
        fast = true;
      }
    } else {
      extra |= W3_RANGE;
    }
  } else if (l >= 0) {
    extra |= W3_RANGE;
  }
  if (fast) {
    // the read planes first: their (coalesced) loads are in flight while the tile loads are issued and waited for
    uint32_t rlo[NP], rhi[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      const bool have = k < rv.n_words;
      rlo[k] = have ? ldg32(rv.rlo + (int64_t)k * rv.stride + i * rv.pair_stride) : 0u;
      rhi[k] = have ? ldg32(rv.rhi + (int64_t)k * rv.stride + i * rv.pair_stride) : 0u;
    }
    scan_fast<NP, T>(g, cfg, ga, gb, l, minus_span, read_n, rlo, rhi, rv, i, best);
  }
  finish(best, p.a_start, p.b_end, l, backsplice, extra, out);
}

// ---------------------------------------------------------------- ASCII -> planes (32 bases per word)
FC_HD void pack32(const uint8_t* src, int count, uint32_t& lo, uint32_t& hi, uint32_t& nn) {
  uint32_t a = 0, b = 0, c = 0;
  for (int j = 0; j < count; ++j) {
    uint32_t ch = src[j] & 0xDFu;  // upper-case
    uint32_t code = 0, isn = 0;
    if (ch == 'A') code = 0;
    else if (ch == 'C') code = 1;
    else if (ch == 'G') code = 2;
    else if (ch == 'T') code = 3;
    else isn = 1;
    a |= (code & 1u) << j;
    b |= (code >> 1) << j;
    c |= isn << j;
  }
  lo = a;
  hi = b;
  nn = c;
}

// tile geometry for windows of up to `w` bases: T sectors per tile, P payload bases; tiles start every TILE_STRIDE
// bases, so a window starting anywhere fits when w <= P - (TILE_STRIDE - 1)
constexpr int TILE_STRIDE = 32;
FC_HD void tile_geometry(int w, int& T, int& P, int& cap) {
  T = w <= 93 ? 1 : (w <= 221 ? 2 : 4);
  P = 128 * T - 4;
  cap = P - (TILE_STRIDE - 1);
  if (cap > 256) cap = 256;
}

}  // namespace fc
