// TEST TOOL (never shipped, never loaded by the product): compiles the DEVICE scan code of
// find_circ2_b200/csrc/scan_core.cuh for the host so that the CPU-only development container can check the
// bit-parallel formulation against the oracle before any GPU time is spent.  The genome arrays are built here
// with the same layout genome.cu produces on the device.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../find_circ2_b200/csrc/scan_core.cuh"

namespace {
constexpr int64_t PAD = 4096, ALIGN = 2048;
struct HostGenome {
  std::vector<uint32_t> seq2, nmask, nsum;
  std::vector<int64_t> off, size;
};
}  // namespace

template <int NW>
static void run(const fc::GenomeView& gv, const fc::ScanCfg& cfg, int64_t n, const int32_t* chrom, const int32_t* a_start,
                const int32_t* b_end, const int32_t* l, const uint8_t* flags, const uint32_t* rd2, const uint32_t* rdn,
                int n_words, fc::HitOut* out, int force_per_base) {
  for (int64_t i = 0; i < n; ++i) {
    fc::PairArgs p{chrom[i], a_start[i], b_end[i], l[i], flags[i]};
    fc::NoEmit ne;
    fc::scan_pair<NW>(gv, cfg, p, rd2, rdn, n, i, n_words, out[i], ne, force_per_base != 0);
  }
}

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
  int64_t total = (off + ALIGN - 1) / ALIGN * ALIGN + ALIGN;
  g->seq2.assign(total / 16 + 64, 0u);
  g->nmask.assign(total / 32 + 64, 0xFFFFFFFFu);
  g->nsum.assign(total / 2048 + 64, 0xFFFFFFFFu);
  for (int i = 0; i < n_chrom; ++i) {
    // clear the summary bits of fully covered blocks, then set them again where an N occurs
    int64_t o = g->off[i];
    for (int64_t b = 0; b * 64 < sizes[i]; ++b) {
      bool anyn = false;
      for (int j = 0; j < 64; ++j) {
        int64_t p = b * 64 + j;
        uint32_t code = 0, isn = 1;
        if (p < sizes[i]) {
          uint8_t c = seqs[i][p] & 0xDF;
          if (c == 'A') { code = 0; isn = 0; }
          else if (c == 'C') { code = 1; isn = 0; }
          else if (c == 'G') { code = 2; isn = 0; }
          else if (c == 'T') { code = 3; isn = 0; }
        }
        int64_t gp = o + p;
        g->seq2[gp >> 4] = (g->seq2[gp >> 4] & ~(3u << (2 * (gp & 15)))) | (code << (2 * (gp & 15)));
        if (!isn) g->nmask[gp >> 5] &= ~(1u << (gp & 31));
        anyn |= isn;
      }
      int64_t blk = (o >> 6) + b;
      if (!anyn) g->nsum[blk >> 5] &= ~(1u << (blk & 31));
    }
  }
  return g;
}

void hh_genome_free(void* h) { delete static_cast<HostGenome*>(h); }

// ASCII internal reads [n][stride] -> word-major rd2/rdn, flags |= READ_N
void hh_pack_reads(int64_t n, const uint8_t* ascii, int stride, const int32_t* l, int n_words, uint32_t* rd2,
                   uint32_t* rdn, uint8_t* flags) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t any = 0;
    for (int w = 0; w < n_words; ++w) {
      int count = l[i] - 16 * w;
      count = count < 0 ? 0 : (count > 16 ? 16 : count);
      uint32_t w2 = 0, wn = 0;
      if (count > 0) fc::pack16(ascii + i * (int64_t)stride + 16 * w, count, w2, wn);
      rd2[(int64_t)w * n + i] = w2;
      rdn[(int64_t)w * n + i] = wn;
      any |= wn;
    }
    if (any) flags[i] |= 4;
  }
}

int hh_scan(void* h, int margin, int maxdist, int noncanonical, int strandpref, int nw, int64_t n, const int32_t* chrom,
            const int32_t* a_start, const int32_t* b_end, const int32_t* l, const uint8_t* flags, const uint32_t* rd2,
            const uint32_t* rdn, int n_words, fc::HitOut* out, int force_per_base) {
  HostGenome* g = static_cast<HostGenome*>(h);
  fc::GenomeView gv;
  gv.seq2 = g->seq2.data();
  gv.nmask = g->nmask.data();
  gv.nsum = g->nsum.data();
  gv.chrom_off = g->off.data();
  gv.chrom_size = g->size.data();
  gv.n_chrom = (int32_t)g->off.size();
  gv.pad = (int32_t)PAD;
  fc::ScanCfg cfg{margin, maxdist, noncanonical, strandpref};
  switch (nw) {
    case 3: run<3>(gv, cfg, n, chrom, a_start, b_end, l, flags, rd2, rdn, n_words, out, force_per_base); break;
    case 5: run<5>(gv, cfg, n, chrom, a_start, b_end, l, flags, rd2, rdn, n_words, out, force_per_base); break;
    case 8: run<8>(gv, cfg, n, chrom, a_start, b_end, l, flags, rd2, rdn, n_words, out, force_per_base); break;
    case 12: run<12>(gv, cfg, n, chrom, a_start, b_end, l, flags, rd2, rdn, n_words, out, force_per_base); break;
    case 16: run<16>(gv, cfg, n, chrom, a_start, b_end, l, flags, rd2, rdn, n_words, out, force_per_base); break;
    default: return -1;
  }
  return 0;
}
}
