// fc_internal.cuh -- context layout and helpers shared by the translation units of libfindcirc_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/findcirc_b200.h"
#include "emit_core.cuh"
#include "scan_core.cuh"

struct fc_genome {
  bool loaded = false;
  bool shared = false;  // a view of another context's store (fc_genome_share): never freed or rebuilt here
  std::vector<std::string> names;
  std::vector<int64_t> sizes;
  std::vector<int64_t> offs;  // global base offset of each chromosome
  int64_t total = 0;          // padded length in bases
  int64_t n_bases = 0, n_n = 0, n_other = 0;
  uint32_t* d_plo = nullptr;
  uint32_t* d_phi = nullptr;
  uint32_t* d_pn = nullptr;
  uint32_t* d_tiles = nullptr;
  int tile_T = 0, tile_S = 0, tile_W = 0;
  uint64_t tile_magic = 0;
  int64_t tile_bytes = 0;
  int64_t* d_off = nullptr;
  int64_t* d_size = nullptr;
  int64_t dev_bytes = 0;
  fc::GenomeView view() const {
    fc::GenomeView v;
    v.plo = d_plo;
    v.phi = d_phi;
    v.pn = d_pn;
    v.chrom_off = d_off;
    v.chrom_size = d_size;
    v.n_chrom = (int32_t)names.size();
    v.pad = FC_GENOME_PAD;
    v.tiles = d_tiles;
    v.tile_T = d_tiles ? tile_T : 0;
    v.tile_W = d_tiles ? tile_W : 0;
    return v;
  }
};

// growable device buffer
struct fc_dbuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes, cudaStream_t st, bool keep, size_t used);
  void release();
};

struct fc_agg {
  fc_dbuf recs;         // fc_jrec[n_recs]
  int64_t n_recs = 0;
  bool n_exact = true;  // n_recs is the exact count (false: upper bound, the exact value is in the device counter)
  fc_dbuf junctions;    // fc_junction[n_junc]
  int64_t n_junc = -1;  // -1: not finalized
  uint64_t max_idx = 0; // upper bound of fc_jrec.idx seen so far (~0: unknown)
  uint64_t idx_lo = ~0ull;  // smallest idx_base of the fc_agg_emit calls
  bool range_declared = false;  // [idx_lo, max_idx) was given by fc_agg_set_idx_range and holds for every record
  bool unordered = false;  // records are not in idx order (emit claims slots per CTA, peer-to-peer emit, append)
  unsigned long long* h_pinned = nullptr;  // pinned landing zone of the counters
  // fused emit + exchange over peer memory
  bool p2p_enabled = false;
  bool p2p_local = false;  // peers are contexts of this process on this device (tests): the barrier publishes, never waits
  bool p2p_ipc = false;    // the peer pointers were opened with cudaIpcOpenMemHandle (to be closed)
  bool p2p_open = false;   // records were sent since the last barrier: the owners cannot reduce yet
  int p2p_world = 1, p2p_rank = 0;
  int64_t p2p_capacity = 0;   // records this rank can receive per step (its buffer holds two such halves)
  int64_t p2p_slice_cap = 0;  // records per (parity, source) slice = min capacity / world
  void* p2p_recs[8] = {};
  void* p2p_cnt[8] = {};
  unsigned long long barrier_epoch = 0;  // fc_p2p_barrier calls so far (the same on every rank); its low bit = parity
  double p2p_timeout_s = 60.0;           // how long the barrier kernel waits for a peer before it gives up
  fc_dbuf scratch[8];
  fc_dbuf p2p_compact;  // sort-based path, multi-GPU: the slices of a step as one run
  fc_dbuf cub_tmp;
  fc_dbuf counters;     // small device counters
  fc_dbuf htab[3];      // sort-based path: hash sets for the distinct counts (reads, fragment names)
  // sort-free path: junction-key table and accumulators (kept clean between calls), distinct set
  fc_dbuf f_keys, f_sets, f_acc;
  fc_dbuf f_part, f_pcur;       // partitioned distinct counts: entries, fill counts (kept all-zero between calls)
  bool f_dirty = false;
  int64_t nj_hint = -1;         // junctions of the last batch this context reduced on the sort-free path (-1: none yet)
  cudaStream_t side = nullptr;  // early clear of the distinct set (see clear_sets_early)
  cudaEvent_t ev_side = nullptr, ev_main = nullptr;
  bool timing = false;          // keep the per-stage device times of fc_agg_finalize (fc_agg_set_timing)
  float stage_us[8] = {};       // clear, accumulate, distinct, mark, finish, copy
  bool sets_clean = false;
  size_t sets_used = 0;         // bytes of f_sets that calls have touched (and an early clear covers)
};

struct fc_ctx {
  int device = 0;
  std::string err;
  fc_genome genome;
  fc_agg agg;
  cudaStream_t own_stream = nullptr;  // used by the host-buffer convenience calls
  cudaStream_t own_stream2 = nullptr; // second lane of the chunked host-buffer path
  cudaEvent_t ev_chunk[2] = {nullptr, nullptr};
  fc_dbuf host_path[16];              // staging for fc_scan_host / fc_batch_host
  fc_dbuf tie_off;                    // fc_batch_ties_host: exclusive prefix sums of n_hits
  fc_dbuf pk[4];                      // packed form (fc_batch) of an fc_pairs batch: descriptors, read rows, N rows, q
  int64_t launches = 0;
  int64_t host_chunk = 1 << 20;       // pairs per chunk of the chunked host-buffer call (FC_HOST_CHUNK at fc_ctx_create)
  int sm_count = 148;
  // device state of the last fc_batch_host call (two-step batches)
  int64_t last_n = 0;
  fc_pairs last_pairs = {};
  bool last_has_payload = false;
  bool last_has_idx = false;  // host_path[15] holds explicit stream positions of the last batch
};

int fc_fail(fc_ctx* ctx, int code, const char* fmt, ...);
int fc_genome_ensure_tiles(fc_ctx* ctx, int w, cudaStream_t st);
int fc_agg_reserve_records(fc_ctx* ctx, int64_t extra, cudaStream_t st);  // room for `extra` more records, now
int fc_agg_emit_begin(fc_ctx* ctx, int64_t n, cudaStream_t st, fc::EmitArgs* e);  // room + where the records go
void fc_agg_emit_end(fc_ctx* ctx, int64_t n, uint64_t idx_base, bool explicit_idx);
int fc_agg_p2p_begin(fc_ctx* ctx, fc::P2PView* pv, unsigned long long** overflow);  // the peer view of the context
void fc_agg_p2p_end(fc_ctx* ctx, uint64_t idx_base, int64_t n);

#define FC_CUDA(ctx, call)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fc_fail((ctx), FC_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define FC_LAUNCH_CHECK(ctx)                                                                       \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return fc_fail((ctx), FC_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
