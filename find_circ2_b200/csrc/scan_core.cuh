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
// B200 formulation (one thread per pair, everything in registers):
//   * windows are funnel-shifted out of two/three 128-bit loads of the 2-bit genome;
//   * mismatch flags mA/mB are one XOR + fold per 16 bases; dist(x) = popc(mA below x) + popc(mB at/above x) -> O(l/16);
//   * GT/AG and CT/AC positions are found for all x at once with bit logic on the 2-bit planes, so only the
//     (on average ~1.5) signal-bearing split positions are ever scored;
//   * pairs whose windows touch an N (coarse 64-base summary bit), whose read contains N, that are longer than the
//     compiled register budget, or that run with --non-canonical take an exact per-base path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FC_HD __host__ __device__ __forceinline__
#else
#define FC_HD inline
#endif

namespace fc {

struct GenomeView {
  const uint32_t* seq2;   // 16 bases / word, A0 C1 G2 T3 (N stored as 0)
  const uint32_t* nmask;  // 32 bases / word, 1 = not ACGT
  const uint32_t* nsum;   // one bit per 64-base block: block contains a non-ACGT base
  const int64_t* chrom_off;   // global index of base 0 of each chromosome (multiple of 128, >= PAD)
  const int64_t* chrom_size;
  int32_t n_chrom;
  int32_t pad;  // bases of N padding on both sides of every chromosome
};

struct ScanCfg {
  int32_t margin, maxdist, noncanonical, strandpref;
};

struct HitOut {
  int32_t start, end;
  uint32_t w2, w3;
};

constexpr uint32_t M55 = 0x55555555u;
constexpr uint32_t SIG_GTAG = 2u | (3u << 3) | (0u << 6) | (2u << 9);
constexpr uint32_t W3_RANGE = 1u << 30;
constexpr uint32_t W3_SLOW = 1u << 31;

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
FC_HD uint32_t ldg32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
struct U4 {
  uint32_t x, y, z, w;
};
FC_HD U4 ldg128(const uint32_t* p) {  // p 16-byte aligned
#if defined(__CUDA_ARCH__)
  uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  return U4{v.x, v.y, v.z, v.w};
#else
  return U4{p[0], p[1], p[2], p[3]};
#endif
}

// ---------------------------------------------------------------- running best / tie bookkeeping
struct Best {
  int score;     // best score so far
  int n_ties;    // hits with that score
  int n_total;   // all hits
  int x;         // split position of the first best hit
  uint32_t info; // strand | sig<<1 | dist<<16 | ov<<24 of the first best hit
  FC_HD void init() {
    score = -100000;
    n_ties = 0;
    n_total = 0;
    x = -1;
    info = 0;
  }
  // hits must be offered in the reference's list order (ascending x, '+' before '-')
  FC_HD void offer(int s, int xx, uint32_t strand, uint32_t sig, int dist, int ov) {
    n_total++;
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
  uint32_t nm = ldg32(g.nmask + (gp >> 5));
  if ((nm >> (gp & 31)) & 1u) return 4;
  return (int)((ldg32(g.seq2 + (gp >> 4)) >> (2 * (gp & 15))) & 3u);
}
FC_HD int rcode(const uint32_t* rd2, const uint32_t* rdn, int64_t n, int64_t i, int j, bool has_n) {
  int w = j >> 4, sh = 2 * (j & 15);
  if (has_n && ((ldg32(rdn + (int64_t)w * n + i) >> sh) & 1u)) return 4;
  return (int)((ldg32(rd2 + (int64_t)w * n + i) >> sh) & 3u);
}
FC_HD uint32_t comp_code(uint32_t c) { return c == 4 ? 4u : 3u - c; }

// Emit: functor called for every hit in reference list order; used by the tie enumerator.
struct NoEmit {
  FC_HD void operator()(int, int, uint32_t, uint32_t, int, int) const {}
};

template <class Emit>
FC_HD void scan_per_base(const GenomeView& g, const ScanCfg& cfg, int64_t ga, int64_t gb, int l, bool minus_span,
                         const uint32_t* rd2, const uint32_t* rdn, int64_t n, int64_t i, bool has_n, Best& best,
                         Emit& emit) {
  // dist(0): everything explained by the acceptor side
  int dist = 0;
  for (int j = 0; j < l; ++j) dist += gcode(g, gb + j + 2) != rcode(rd2, rdn, n, i, j, has_n);
  for (int x = 0; x <= l; ++x) {
    if (dist <= cfg.maxdist) {
      uint32_t a0 = gcode(g, ga + x), a1 = gcode(g, ga + x + 1), b0 = gcode(g, gb + x), b1 = gcode(g, gb + x + 1);
      uint32_t sig = a0 | (a1 << 3) | (b0 << 6) | (b1 << 9);
      uint32_t rc = comp_code(b1) | (comp_code(b0) << 3) | (comp_code(a1) << 6) | (comp_code(a0) << 9);
      int ov = anchor_overlap(x, l, cfg.margin);
      int base = -10 * dist - ov;
      bool is_gtag = sig == SIG_GTAG, is_ctac = rc == SIG_GTAG;
      if (cfg.noncanonical) {
        int sp = base + (is_gtag ? 20 : 0) + ((cfg.strandpref && !minus_span) ? 100 : 0);
        int sm = base + (is_ctac ? 20 : 0) + ((cfg.strandpref && minus_span) ? 100 : 0);
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
      int r = rcode(rd2, rdn, n, i, x, has_n);
      dist += (gcode(g, ga + x) != r) - (gcode(g, gb + x + 2) != r);
    }
  }
}

// ---------------------------------------------------------------- bit-parallel path
// Extract NW 32-bit words (16 bases each) starting at global base index gp.
template <int NW>
FC_HD void load_window(const GenomeView& g, int64_t gp, int nbases, uint32_t (&W)[NW + 1]) {
  constexpr int NQ = (NW + 4 + 3) / 4;  // 128-bit loads that can be touched: NW words + up to 3 words of misalignment + 1
  constexpr int NS = NQ * 4;
  uint32_t s[NS + 2];
  const int64_t q0 = gp >> 6;        // 64 bases per 16 bytes
  const int o = (int)(gp & 63);      // base offset inside the first 16-byte chunk
  const uint32_t* base = g.seq2 + q0 * 4;
  const int last_q = (o + nbases - 1) >> 6;  // index of the last chunk actually needed
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if (q <= last_q) {
      U4 v = ldg128(base + q * 4);
      s[q * 4 + 0] = v.x;
      s[q * 4 + 1] = v.y;
      s[q * 4 + 2] = v.z;
      s[q * 4 + 3] = v.w;
    } else {
      s[q * 4 + 0] = s[q * 4 + 1] = s[q * 4 + 2] = s[q * 4 + 3] = 0;
    }
  }
  s[NS] = 0;
  s[NS + 1] = 0;
  const int wo = o >> 4;                  // whole-word offset 0..3
  const uint32_t bo = (uint32_t)(o & 15) * 2;  // bit offset 0..30
  if (wo & 2) {
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] = s[k + 2];
  }
  if (wo & 1) {
#pragma unroll
    for (int k = 0; k < NS; ++k) s[k] = s[k + 1];
  }
#pragma unroll
  for (int k = 0; k < NW; ++k) W[k] = funnel_r(s[k], s[k + 1], bo);
  W[NW] = 0;
}

// true when any 64-base block overlapping [gp, gp+nbases) contains a non-ACGT base
FC_HD bool window_has_n(const GenomeView& g, int64_t gp, int nbases) {
  int64_t b0 = gp >> 6, b1 = (gp + nbases - 1) >> 6;
  int64_t w0 = b0 >> 5, w1 = b1 >> 5;
  uint32_t lo = ldg32(g.nsum + w0);
  uint32_t hi = (w1 != w0) ? ldg32(g.nsum + w0 + 1) : 0u;
  uint64_t bits = (((uint64_t)hi << 32) | lo) >> (b0 & 31);
  int nb = (int)(b1 - b0 + 1);  // <= 32 for windows up to ~2000 bases
  uint64_t mask = nb >= 64 ? ~0ull : ((1ull << nb) - 1ull);
  return (bits & mask) != 0ull;
}

FC_HD uint32_t valid_mask(int nbases, int word) {  // even bits of the bases < nbases that live in `word`
  int k = nbases - 16 * word;
  if (k <= 0) return 0u;
  if (k >= 16) return M55;
  return ((1u << (2 * k)) - 1u) & M55;
}

// canonical-signal scan with NW window words in registers (covers l + 2 <= 16*NW)
template <int NW>
FC_HD void scan_bits(const GenomeView& g, const ScanCfg& cfg, int64_t ga, int64_t gb, int l, bool minus_span,
                     const uint32_t* rd2, int64_t n, int64_t i, int n_words, Best& best) {
  uint32_t A[NW + 1], B[NW + 1];
  load_window<NW>(g, ga, l + 2, A);
  load_window<NW>(g, gb, l + 2, B);

  // splice signal at every split position x (bit 2*(x%16) of word x/16):
  //   GT..AG: A[x]=G A[x+1]=T B[x]=A B[x+1]=G ;  CT..AC: A[x]=C A[x+1]=T B[x]=A B[x+1]=C     (find_circ.py:924-954)
  uint32_t sigP[NW], sigM[NW];
  uint32_t any = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    uint32_t a = A[k], a1 = (A[k] >> 2) | (A[k + 1] << 30);
    uint32_t b = B[k], b1 = (B[k] >> 2) | (B[k + 1] << 30);
    uint32_t a_lo = a, a_hi = a >> 1, a1_lo = a1, a1_hi = a1 >> 1;
    uint32_t b_lo = b, b_hi = b >> 1, b1_lo = b1, b1_hi = b1 >> 1;
    uint32_t common = (a1_hi & a1_lo) & ~(b_hi | b_lo);          // A[x+1]==T and B[x]==A
    uint32_t gg = (a_hi & ~a_lo) & (b1_hi & ~b1_lo);             // A[x]==G and B[x+1]==G
    uint32_t cc = (~a_hi & a_lo) & (~b1_hi & b1_lo);             // A[x]==C and B[x+1]==C
    uint32_t vm = valid_mask(l + 1, k);
    sigP[k] = common & gg & vm;
    sigM[k] = common & cc & vm;
    any |= sigP[k] | sigM[k];
  }
  if (!any) return;  // no split position carries a canonical signal (most decoy pairs end here)

  // mismatch flags of the read against the donor window (A[i] vs R[i]) and the acceptor window (B[i+2] vs R[i])
  uint32_t mA[NW], mB[NW];
  int totalB = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    uint32_t r = (k < n_words) ? ldg32(rd2 + (int64_t)k * n + i) : 0u;
    uint32_t b2 = (B[k] >> 4) | (B[k + 1] << 28);
    uint32_t xa = A[k] ^ r, xb = b2 ^ r;
    uint32_t vm = valid_mask(l, k);
    mA[k] = (xa | (xa >> 1)) & vm;
    mB[k] = (xb | (xb >> 1)) & vm;
    totalB += popc32(mB[k]);
  }

  int accA = 0, accB = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    uint32_t c = sigP[k] | sigM[k];
    while (c) {
      int bit = ffs32(c);
      c &= c - 1;
      uint32_t below = (1u << bit) - 1u;
      int dist = accA + popc32(mA[k] & below) + (totalB - accB - popc32(mB[k] & below));
      if (dist <= cfg.maxdist) {
        int x = 16 * k + (bit >> 1);
        uint32_t strand = (sigM[k] >> bit) & 1u;
        int ov = anchor_overlap(x, l, cfg.margin);
        int s = 20 - 10 * dist - ov;
        if (cfg.strandpref && ((strand != 0u) == minus_span)) s += 100;
        best.offer(s, x, strand, SIG_GTAG, dist, ov);
      }
    }
    accA += popc32(mA[k]);
    accB += popc32(mB[k]);
  }
}

// ---------------------------------------------------------------- one pair
struct PairArgs {
  int32_t chrom, a_start, b_end, l;
  uint32_t flags;  // FC_PF_*
};

template <int NW, class Emit>
FC_HD void scan_pair(const GenomeView& g, const ScanCfg& cfg, const PairArgs& p, const uint32_t* rd2,
                     const uint32_t* rdn, int64_t n, int64_t i, int n_words, HitOut& out, Emit& emit,
                     bool force_per_base) {
  Best best;
  best.init();
  const bool backsplice = p.flags & 1u, minus_span = p.flags & 2u, read_n = p.flags & 4u;
  const int l = p.l;
  uint32_t extra = 0;
  if (l >= 0 && p.chrom >= 0 && p.chrom < g.n_chrom) {
    const int64_t size = g.chrom_size[p.chrom];
    const int64_t wa0 = p.a_start, wa1 = (int64_t)p.a_start + l + 2;
    const int64_t wb1 = p.b_end, wb0 = (int64_t)p.b_end - (l + 2);
    // windows must overlap the chromosome (find_circ.py:194-211 pads the overhang with N; a window entirely outside
    // is undefined there) and the overhang must fit in the padding
    // (a window that merely touches the boundary still has the right length there and reads all-N)
    const bool ok = wa0 <= size && wa1 >= 0 && wb0 <= size && wb1 >= 0 && wa0 >= -(int64_t)g.pad &&
                    wa1 <= size + g.pad && wb0 >= -(int64_t)g.pad && wb1 <= size + g.pad;
    if (ok) {
      const int64_t off = g.chrom_off[p.chrom];
      const int64_t ga = off + wa0, gb = off + wb0;
      bool slow = force_per_base || cfg.noncanonical || read_n || (l + 2 > 16 * NW);
      if (!slow) slow = window_has_n(g, ga, l + 2) || window_has_n(g, gb, l + 2);
      if (slow) {
        extra |= W3_SLOW;
        scan_per_base(g, cfg, ga, gb, l, minus_span, rd2, rdn, n, i, read_n, best, emit);
      } else {
        scan_bits<NW>(g, cfg, ga, gb, l, minus_span, rd2, n, i, n_words, best);
      }
    } else {
      extra |= W3_RANGE;
    }
  } else if (l >= 0) {
    extra |= W3_RANGE;
  }
  finish(best, p.a_start, p.b_end, l, backsplice, extra, out);
}

// ---------------------------------------------------------------- ASCII -> 2-bit
FC_HD void pack16(const uint8_t* src, int count, uint32_t& w2, uint32_t& wn) {
  uint32_t a = 0, nn = 0;
  for (int j = 0; j < count; ++j) {
    uint32_t c = src[j] & 0xDFu;  // upper-case
    uint32_t code, isn = 0;
    if (c == 'A') code = 0;
    else if (c == 'C') code = 1;
    else if (c == 'G') code = 2;
    else if (c == 'T') code = 3;
    else { code = 0; isn = 1; }
    a |= code << (2 * j);
    nn |= isn << (2 * j);
  }
  w2 = a;
  wn = nn;
}

}  // namespace fc
