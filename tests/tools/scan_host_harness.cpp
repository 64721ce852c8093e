// TEST TOOL (never shipped, never loaded by the product): compiles the DEVICE scan code of
// find_circ2_b200/csrc/scan_core.cuh for the host so that the CPU-only development container can check the
// bit-parallel formulation against the oracle before any GPU time is spent.  The genome planes and tiles are built
// here with the same layout genome.cu produces on the device.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../find_circ2_b200/csrc/scan_core.cuh"

namespace {
constexpr int64_t PAD = 4096, ALIGN = 2048, SLACK = 4096;
struct HostGenome {
  std::vector<uint32_t> plo, phi, pn, tiles;
  std::vector<int64_t> off, size;
  int64_t total = 0;
  int T = 0, S = 0, W = 0;
  uint64_t magic = 0;
};

uint32_t extract32(const std::vector<uint32_t>& plane, int64_t bit) {
  const uint32_t* p = plane.data() + (bit >> 5);
  return fc::funnel_r(p[0], p[1], (uint32_t)(bit & 31));
}

template <int NP, int T>
void run(const fc::GenomeView& gv, const fc::ScanCfg& cfg, const fc::ReadView& rv, const int32_t* chrom,
         const int32_t* a_start, const int32_t* b_end, const int32_t* l, const uint8_t* flags, fc::HitOut* out,
         int force_per_base) {
  for (int64_t i = 0; i < rv.n; ++i) {
    fc::PairArgs p{chrom[i], a_start[i], b_end[i], l[i], flags[i]};
    fc::NoEmit ne;
    fc::scan_pair<NP, T>(gv, cfg, p, rv, i, out[i], ne, force_per_base != 0);
  }
}
}  // namespace

extern "C" {

void* hh_genome_build(int n_chrom, const uint8_t* const* seqs, const int64_t* sizes) {
  HostGenome* g = new HostGenome();
  int64_t off = PAD;
  for (int i = 0; i < n_chrom; ++i) {
    off = (off + ALIGN - 1) / ALIGN * ALIGN;
    g->off.push_back(off);
    g->size.push_back(sizes[i]);
    off += sizes[i] + PAD;
  }
  g->total = (off + ALIGN - 1) / ALIGN * ALIGN + ALIGN;
  size_t words = (size_t)((g->total + SLACK) / 32 + 64);
  g->plo.assign(words, 0u);
  g->phi.assign(words, 0u);
  g->pn.assign(words, 0xFFFFFFFFu);
  for (int i = 0; i < n_chrom; ++i) {
    for (int64_t p = 0; p < sizes[i]; ++p) {
      uint8_t c = seqs[i][p] & 0xDF;
      uint32_t code = 0, isn = 1;
      if (c == 'A') { code = 0; isn = 0; }
      else if (c == 'C') { code = 1; isn = 0; }
      else if (c == 'G') { code = 2; isn = 0; }
      else if (c == 'T') { code = 3; isn = 0; }
      int64_t gp = g->off[i] + p;
      uint32_t bit = 1u << (gp & 31);
      if (code & 1u) g->plo[gp >> 5] |= bit;
      if (code & 2u) g->phi[gp >> 5] |= bit;
      if (!isn) g->pn[gp >> 5] &= ~bit;
    }
  }
  return g;
}

// same arithmetic as build_tiles_kernel in genome.cu
void hh_build_tiles(void* h, int w) {
  HostGenome* g = static_cast<HostGenome*>(h);
  if (w < 8) w = 8;
  if (w > 256) w = 256;
  int T, P, cap;
  fc::tile_geometry(w, T, P, cap);
  const int S = fc::TILE_STRIDE;
  int64_t n_tiles = g->total / S + 2;
  g->tiles.assign((size_t)n_tiles * 8 * T + 64, 0u);
  const int PW = 4 * T;
  for (int64_t t = 0; t < n_tiles; ++t) {
    int64_t g0 = t * (int64_t)S;
    uint32_t* dst = g->tiles.data() + t * (int64_t)(8 * T);
    uint32_t anyn = 0;
    for (int j = 0; j < PW; ++j) {
      uint32_t lo = extract32(g->plo, g0 + 32 * j), hi = extract32(g->phi, g0 + 32 * j), nn = extract32(g->pn, g0 + 32 * j);
      if (j == PW - 1) {
        lo &= 0x0FFFFFFFu;
        hi &= 0x0FFFFFFFu;
        nn &= 0x0FFFFFFFu;
      }
      anyn |= nn;
      if (j == PW - 1 && anyn) lo |= fc::TILE_FLAG_N;
      dst[j] = lo;
      dst[PW + j] = hi;
    }
  }
  g->T = T;
  g->S = S;
  g->W = cap;
}

void hh_genome_free(void* h) { delete static_cast<HostGenome*>(h); }

// ASCII internal reads [n][stride] -> word-major planes, flags |= READ_N
void hh_pack_reads(int64_t n, const uint8_t* ascii, int stride, const int32_t* l, int n_words, uint32_t* rlo,
                   uint32_t* rhi, uint32_t* rn, uint8_t* flags) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t any = 0;
    for (int w = 0; w < n_words; ++w) {
      int count = l[i] - 32 * w;
      count = count < 0 ? 0 : (count > 32 ? 32 : count);
      uint32_t lo = 0, hi = 0, nn = 0;
      if (count > 0) fc::pack32(ascii + i * (int64_t)stride + 32 * w, count, lo, hi, nn);
      rlo[(int64_t)w * n + i] = lo;
      rhi[(int64_t)w * n + i] = hi;
      rn[(int64_t)w * n + i] = nn;
      any |= nn;
    }
    if (any) flags[i] |= 4;
  }
}

// mode: 0 = kernel choice as fc_scan makes it (tiles), 1 = no tile store (master planes), 2 = force per-base
int hh_scan(void* h, int margin, int maxdist, int noncanonical, int strandpref, int mode, int max_l, int64_t n,
            const int32_t* chrom, const int32_t* a_start, const int32_t* b_end, const int32_t* l, const uint8_t* flags,
            const uint32_t* rlo, const uint32_t* rhi, const uint32_t* rn, int n_words, fc::HitOut* out) {
  HostGenome* g = static_cast<HostGenome*>(h);
  fc::GenomeView gv;
  gv.plo = g->plo.data();
  gv.phi = g->phi.data();
  gv.pn = g->pn.data();
  gv.chrom_off = g->off.data();
  gv.chrom_size = g->size.data();
  gv.n_chrom = (int32_t)g->off.size();
  gv.pad = (int32_t)PAD;
  const bool use_tiles = mode == 0 && !noncanonical && !g->tiles.empty();
  gv.tiles = use_tiles ? g->tiles.data() : nullptr;
  gv.tile_T = use_tiles ? g->T : 0;
  gv.tile_W = use_tiles ? g->W : 0;
  fc::ScanCfg cfg{margin, maxdist, noncanonical, strandpref};
  fc::ReadView rv{rlo, rhi, rn, n, n_words, n};
  const int need = max_l + 2;
  const int force = mode == 2;
  switch (gv.tile_T) {
    case 1:
      if (need <= 64) run<2, 1>(gv, cfg, rv, chrom, a_start, b_end, l, flags, out, force);
      else run<3, 1>(gv, cfg, rv, chrom, a_start, b_end, l, flags, out, force);
      break;
    case 2:
      if (need <= 128) run<4, 2>(gv, cfg, rv, chrom, a_start, b_end, l, flags, out, force);
      else run<8, 2>(gv, cfg, rv, chrom, a_start, b_end, l, flags, out, force);
      break;
    case 4: run<8, 4>(gv, cfg, rv, chrom, a_start, b_end, l, flags, out, force); break;
    default: run<8, 0>(gv, cfg, rv, chrom, a_start, b_end, l, flags, out, force); break;
  }
  return 0;
}
}
