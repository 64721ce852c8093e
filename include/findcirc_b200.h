/*
 * findcirc_b200.h -- C ABI of libfindcirc_b200.so: the B200-native (sm_100a) breakpoint scan and junction
 * aggregation behind find_circ.py's command line.
 *
 * The reference (feiyue126/find_circ2, /root/reference/find_circ.py v1.99) is one Python process with no
 * FFI seam; the drop-in boundary is its internal call structure.  Each entry point
 * below names the reference code it replaces:
 *
 *   genome store    Track / GenomeAccessor / indexed_fasta.get_data      find_circ.py:103-215, 242-371
 *   scan            JunctionSpan.find_breakpoints + Splice.score          find_circ.py:766-806, 854-974
 *   aggregation     SpliceSiteStorage.add + Hit.add + Hit reductions      find_circ.py:486-600, 657-690
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success and a negative FC_E_* code
 * on failure (fc_last_error() gives the text); nothing throws or exits across the boundary; one calling
 * thread per context; a context is bound to one CUDA device.  Pointers named d_* are DEVICE pointers
 * (e.g. torch tensors' data_ptr()), h_* are host pointers.  `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream); device-pointer calls are asynchronous on that stream.
 *
 * There is no CPU implementation behind these entry points.
 */
#ifndef FINDCIRC_B200_H
#define FINDCIRC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FC_ABI_VERSION 2

/* error codes */
#define FC_OK 0
#define FC_E_CUDA -1      /* a CUDA runtime call failed */
#define FC_E_ARG -2       /* bad argument */
#define FC_E_IO -3        /* file could not be read */
#define FC_E_FORMAT -4    /* malformed FASTA */
#define FC_E_NOGENOME -5  /* scan before a genome was loaded */
#define FC_E_RANGE -6     /* an anchor window lies entirely outside its chromosome (undefined in the reference) */
#define FC_E_NOMEM -7
#define FC_E_STATE -8     /* call order violated */
#define FC_E_COLLISION -9 /* 64-bit junction-key hash collision survived all reseeds */

typedef struct fc_ctx fc_ctx;

/* Scan switches -- the subset of find_circ.py's optparse table (find_circ.py:383-413) that reaches
 * find_breakpoints():  -a/--anchor, -m/--margin, -d/--max-mismatch, --non-canonical, --strand-pref. */
typedef struct fc_scan_params {
  int32_t asize;        /* find_circ.py:394 */
  int32_t margin;       /* find_circ.py:395 */
  int32_t maxdist;      /* find_circ.py:396 */
  int32_t noncanonical; /* find_circ.py:401 */
  int32_t strandpref;   /* find_circ.py:404 */
  int32_t reserved[3];
} fc_scan_params;

/* ---- per-pair flags (fc_pairs.flags) ---- */
#define FC_PF_BACKSPLICE 1u  /* JunctionSpan.is_backsplice, find_circ.py:850-852 */
#define FC_PF_MINUS 2u       /* JunctionSpan.strand == '-', find_circ.py:834-837 */
#define FC_PF_READ_N 4u      /* the internal read part contains a non-ACGT letter (set by fc_pack_reads) */
#define FC_PF_LINEAR_KIND 8u /* informational: not a back-splice (== !BACKSPLICE) */

/* Struct-of-arrays batch of anchor pairs (one row per JunctionSpan, find_circ.py:821-852), device resident.
 *   a_start = A.pos + (asize - margin)       genomic position of A_flank[0]      (find_circ.py:901)
 *   b_end   = B.aend - (asize - margin)      one past the last base of B_flank   (find_circ.py:902)
 *   l       = len(read_part) - 2*(asize-margin)   number of internal bases        (find_circ.py:904); may be < 0
 *   rlo,rhi low / high bit plane of the 2-bit codes (A0 C1 G2 T3, N stored as 0) of the internal read bases,
 *           word-major: word w of pair i at rlo[w*n + i], base j in word j/32 at bit j%32
 *   rn      same layout, bit set when base j is not A/C/G/T
 *   n_words = words per pair and plane  (>= ceil(max l / 32))
 */
typedef struct fc_pairs {
  int64_t n;
  const int32_t* d_chrom;
  const int32_t* d_a_start;
  const int32_t* d_b_end;
  const int32_t* d_l;
  const uint8_t* d_flags; /* FC_PF_* ; READ_N may be OR-ed in by fc_pack_reads, hence also written */
  const uint32_t* d_rlo;
  const uint32_t* d_rhi;
  const uint32_t* d_rn;
  int32_t n_words;
  int32_t max_l; /* upper bound of l over the batch (selects the kernel specialisation) */
  int64_t plane_stride; /* words between consecutive words of one pair in rlo/rhi/rn; 0 = n */
} fc_pairs;

/* The same batch in the layout the scan kernels read (fc_pairs batches are converted on the device with fc_batch_pack; a
 * native ingest writes it directly).  One 16-byte descriptor per pair, four 32-bit words:
 *   w0  low 32 bits of ga = genome coordinate of A_flank[0]  (fc_genome_chrom_offset(chrom) + a_start)
 *   w1  low 32 bits of gb = genome coordinate of B_flank[0]  (fc_genome_chrom_offset(chrom) + b_end - (l + 2))
 *   w2  bits 0-5 ga >> 32, 6-11 gb >> 32, 12-23 l, 24-26 FC_PF_BACKSPLICE / MINUS / READ_N, 27 FC_META_INVALID (a window
 *       lies outside its chromosome, the chromosome is unknown or l < 0: no hit; ga, gb, l are then 0),
 *       28-29 / 30-31 rows of the same fragment before / after this one (3,3 = unknown).  The rows of a fragment are
 *       neighbours; they lie in one batch, or -- batches that continue one another with idx = idx_base + row -- the batch
 *       boundary that cuts the fragment falls on a multiple of 32 rows
 *   w3  bits 0-23 chromosome id, 24-31 weight denominator (find_circ.py:1084)
 * and the internal read part as one row per pair: n_words words of the lo plane, then n_words words of the hi plane
 * (d_reads), n_words words of the N plane (d_rn; read only for pairs flagged FC_PF_READ_N, may be NULL when none is).
 * n_words = fc_batch_words(max_l): ceil(max_l / 32) rounded up to 1, 2, 4 or 8 (rows are loaded with one vector load). */
#define FC_META_INVALID 8u
#define FC_META_NEGATIVE 16u /* internal to fc_batch_pack: never stored */
typedef struct fc_batch {
  int64_t n;
  const void* d_meta;
  const uint32_t* d_reads;
  const uint32_t* d_rn;
  int32_t n_words;
  int32_t max_l;
} fc_batch;

/* Result of the scan for one pair: the FIRST best-scoring breakpoint (ties keep ascending split position,
 * '+' before '-', find_circ.py:966-974) and the number of ties.  16 bytes, one coalesced store per pair.
 *   start,end  BED coordinates after the back-splice / linear correction (find_circ.py:929-945); valid iff n_hits>0
 *   w2         bits 0-15 n_hits (= number of ties, find_circ.py:969), 16-23 dist, 24-31 anchor overlap
 *   w3         bit 0 strand ('-' = 1), bits 1-12 signal (4 letters x 3 bits, A0 C1 G2 T3 N4, first letter lowest),
 *              bits 13-22 best score + 512, bit 30 window-out-of-range, bit 31 slow (per-base) path was taken
 */
typedef struct fc_hit {
  int32_t start;
  int32_t end;
  uint32_t w2;
  uint32_t w3;
} fc_hit;

/* 48-byte junction record: one per Hit.add() call (find_circ.py:526-582); this is also the unit exchanged
 * between GPUs (hash-partitioned by key). */
#define FC_SK_NAME_KNOWN 8u
#define FC_SK_NAME_DUP 16u
typedef struct fc_jrec {
  uint32_t chrom;
  uint32_t start;
  uint32_t end;
  uint32_t sk;         /* bit0 strand '-', bit1 kind (1 = linear table), bit2 read is its own reverse complement,
                          bit3 FC_SK_NAME_KNOWN: bit4 (FC_SK_NAME_DUP) tells whether an earlier record of the same fragment
                          already supports this junction, qname_hash need not be consulted (set by the scan kernels from the
                          fragment fields of fc_batch descriptors; valid only when a read name occurs in ONE fragment),
                          bits 8-15 weight denominator (weight = 1/den, find_circ.py:1084), bits 16-27 signal */
  uint64_t idx;        /* position in the input stream (orders names and float sums, find_circ.py:684-686, 544) */
  uint64_t read_hash;  /* strand-invariant hash of primary.seq (n_uniq, find_circ.py:581-590) */
  uint64_t qname_hash; /* hash of primary.qname (n_frags, find_circ.py:584-586) */
  int16_t q_left;      /* AS-XS of the genome-left anchor, XS default 0 (find_circ.py:552-559) */
  int16_t q_right;
  uint16_t n_hits;
  uint8_t dist;
  uint8_t ov;
} fc_jrec;

/* 64-byte aggregated junction (one Hit): everything store_list() prints that is not text (find_circ.py:722-730) */
typedef struct fc_junction {
  uint32_t chrom;
  uint32_t start;
  uint32_t end;
  uint32_t sk;         /* bit0 strand, bit1 kind, bits 16-27 signal */
  uint64_t first_idx;  /* smallest idx -> discovery order -> name number */
  double n_weighted;   /* sum of weights in stream order */
  double n_uniq_bridges;
  uint32_t n_spanned;
  uint32_t n_frags;    /* distinct qnames */
  uint32_t n_uniq;     /* len(uniq)/2: distinct read sequences modulo reverse complement */
  int16_t best_q_left;
  int16_t best_q_right;
  uint16_t min_n_hits;
  uint8_t min_dist;
  uint8_t min_ov;
  uint32_t pad;
} fc_junction;

/* ------------------------------------------------------------------ context */
int fc_abi_version(void);
int fc_ctx_create(int device, fc_ctx** out);
void fc_ctx_destroy(fc_ctx* ctx);
const char* fc_last_error(fc_ctx* ctx); /* ctx may be NULL: last error of a failed fc_ctx_create */

/* ------------------------------------------------------------------ genome store
 * Replaces indexed_fasta / GenomeAccessor (find_circ.py:103-215, 329-371).  Chromosomes are packed into device bit
 * planes (2 bits + an N bit per base) in one coordinate space with >= FC_GENOME_PAD bases of 'N' padding around each, so reads outside [0,size) return 'N' as
 * find_circ.py:194-211 does; soft-masked (lower-case) letters are upper-cased as the callers do (:901-902);
 * letters other than ACGTN are stored as N and counted (fc_genome_stats). */
#define FC_GENOME_PAD 4096
int fc_genome_load_fasta(fc_ctx* ctx, const char* path);
/* n_chrom sequences given as ASCII in host memory (synthetic genomes; avoids a FASTA round trip) */
int fc_genome_load_ascii(fc_ctx* ctx, int32_t n_chrom, const char* const* names, const uint8_t* const* seqs,
                         const int64_t* sizes);
/* dst looks at the device store of src (another context of this process on the same device) instead of loading its own
 * copy; src must outlive dst, and reads longer than the store's current tile class must be scanned on src first */
int fc_genome_share(fc_ctx* dst, fc_ctx* src);
int fc_genome_n_chrom(fc_ctx* ctx);
int fc_genome_chrom_name(fc_ctx* ctx, int32_t i, char* buf, int32_t cap);
int64_t fc_genome_chrom_size(fc_ctx* ctx, int32_t i);
int64_t fc_genome_chrom_offset(fc_ctx* ctx, int32_t i); /* genome coordinate of base 0 of chromosome i (fc_batch descriptors) */
int fc_genome_chrom_id(fc_ctx* ctx, const char* name); /* -1 unknown (the reference raises KeyError, :193) */
/* stats[0]=total bases, [1]=N bases, [2]=other non-ACGT letters stored as N, [3]=device bytes */
int fc_genome_stats(fc_ctx* ctx, int64_t stats[4]);
/* genome.get(chrom,start,end,'+').upper() decoded FROM THE DEVICE store into h_out (end-start bytes) */
int fc_genome_fetch(fc_ctx* ctx, int32_t chrom, int64_t start, int64_t end, char* h_out);

/* ------------------------------------------------------------------ read packing
 * d_ascii: n rows of `stride` bytes, row i holds the l[i] internal read bases (read_part[eff:-eff], find_circ.py:895),
 * any case.  Writes rlo / rhi / rn (n_words words per pair and plane, word-major) and ORs FC_PF_READ_N into d_flags. */
int fc_pack_reads(fc_ctx* ctx, int64_t n, const uint8_t* d_ascii, int32_t stride, const int32_t* d_l,
                  int32_t n_words, uint32_t* d_rlo, uint32_t* d_rhi, uint32_t* d_rn, uint8_t* d_flags, void* stream);

/* ------------------------------------------------------------------ breakpoint scan
 * One thread per anchor pair; see find_circ2_b200/csrc/scan_core.cuh. */
int fc_scan(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pairs, fc_hit* d_out, void* stream);
/* --all-hits (find_circ.py:1312-1317): d_tie_off[i] = exclusive prefix sum of n_hits over the pairs (n+1 entries);
 * writes every tie of every pair in rank order to d_ties[d_tie_off[i] ...] */
/* fc_scan and fc_agg_emit in one kernel: every pair that found a breakpoint becomes a junction record (fc_jrec) of the
 * context on the way, d_out still receives one fc_hit per pair.  Same records as fc_scan followed by fc_agg_emit(d_mask =
 * NULL) (find_circ.py:854-974 then :1312-1317, 526-582); d_idx may be NULL (idx = idx_base + pair index). */
int fc_scan_emit(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pairs, fc_hit* d_out, const uint8_t* d_wden,
                 const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash, const uint64_t* d_qname_hash,
                 uint64_t idx_base, const uint64_t* d_idx, void* stream);
int fc_scan_ties(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pairs, const fc_hit* d_hits,
                 const int64_t* d_tie_off, fc_hit* d_ties, void* stream);
/* packed batches: fc_batch_pack converts an fc_pairs batch (d_wden NULL = 1; d_frag NULL = nothing known about fragments,
 * else bits 0-1 / 2-3 of a byte per pair = rows of the same fragment before / after; d_q receives q_a | q_b << 16 when
 * d_q_a / d_q_b are given); fc_scan_batch = fc_scan, fc_scan_emit_batch = fc_scan_emit -- or fc_scan_emit_p2p when the
 * context is connected to peers (fc_p2p_connect) */
int32_t fc_batch_words(int32_t max_l);
int fc_batch_pack(fc_ctx* ctx, const fc_pairs* pairs, const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b,
                  const uint8_t* d_frag, void* d_meta, uint32_t* d_reads, uint32_t* d_rn, uint32_t* d_q, void* stream);
int fc_scan_batch(fc_ctx* ctx, const fc_scan_params* p, const fc_batch* batch, fc_hit* d_out, void* stream);
int fc_scan_emit_batch(fc_ctx* ctx, const fc_scan_params* p, const fc_batch* batch, fc_hit* d_out, const uint32_t* d_q,
                       const uint64_t* d_read_hash, const uint64_t* d_qname_hash, uint64_t idx_base, const uint64_t* d_idx,
                       void* stream);
/* host-buffer convenience (the reference-facing call): copies the batch in, scans, copies results out.
 * h_ascii rows hold the internal read bases.  Pinned host memory makes the copies asynchronous. */
int fc_scan_host(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom, const int32_t* h_a_start,
                 const int32_t* h_b_end, const int32_t* h_l, const uint8_t* h_flags, const uint8_t* h_ascii,
                 int32_t stride, fc_hit* h_out);

/* One batch through the whole device path with HOST buffers: upload scan inputs (+ the aggregation payload when
 * emit != 0), pack, scan, append the first tie of every pair with n_hits > 0 to the device aggregator
 * (== circ_splices.add / linear_splices.add, find_circ.py:1312-1317, 1364-1378), download the hits (h_out may be NULL).
 * Record order (fc_jrec.idx) is idx_base + row. */
int fc_batch_host(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom, const int32_t* h_a_start,
                  const int32_t* h_b_end, const int32_t* h_l, const uint8_t* h_flags, const uint8_t* h_ascii,
                  int32_t stride, const uint8_t* h_wden, const int16_t* h_q_a, const int16_t* h_q_b,
                  const uint64_t* h_read_hash, const uint64_t* h_qname_hash, uint64_t idx_base, int32_t emit,
                  fc_hit* h_out);

/* fc_batch_host with an explicit stream position per row (h_idx, may be NULL = idx_base + row) */
int fc_batch_host_idx(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom, const int32_t* h_a_start,
                      const int32_t* h_b_end, const int32_t* h_l, const uint8_t* h_flags, const uint8_t* h_ascii,
                      int32_t stride, const uint8_t* h_wden, const int16_t* h_q_a, const int16_t* h_q_b,
                      const uint64_t* h_read_hash, const uint64_t* h_qname_hash, const uint64_t* h_idx, uint64_t idx_base,
                      int32_t emit, fc_hit* h_out);

/* Same as fc_batch_host for rows whose internal read part is already packed as bit planes (what fc_ingest_parse emits):
 * h_rlo/h_rhi/h_rn are word-major with `plane_stride` words between the words of one pair; h_idx (may be NULL) gives an
 * explicit stream position per row instead of idx_base + row. */
int fc_batch_host_planes(fc_ctx* ctx, const fc_scan_params* p, int64_t n, const int32_t* h_chrom, const int32_t* h_a_start,
                         const int32_t* h_b_end, const int32_t* h_l, const uint8_t* h_flags, const uint32_t* h_rlo,
                         const uint32_t* h_rhi, const uint32_t* h_rn, int32_t n_words, int64_t plane_stride, int32_t max_l,
                         const uint8_t* h_wden, const int16_t* h_q_a, const int16_t* h_q_b, const uint64_t* h_read_hash,
                         const uint64_t* h_qname_hash, const uint64_t* h_idx, uint64_t idx_base, int32_t emit, fc_hit* h_out);

/* ---- streamed host batches: the main loop of the reference (find_circ.py:1535-1574: read, scan, record, one fragment at a
 * time) as a pipeline of batches.  A stream owns n_slots device-side slots (own CUDA stream each) for batches of up to
 * cap_rows rows x max_words read words; fc_stream_submit queues, without waiting, the copy of one batch in fc_batch layout
 * from HOST memory (pinned -- fc_pinned_alloc -- for the copies to overlap), the scan, the recording of the pairs with a
 * breakpoint when emit != 0 (into this context, or into the owner ranks' buffers when the context is connected to peers)
 * and the copy of the results back; fc_stream_wait blocks until the slot's results are in host memory.  The caller owns all
 * host arrays between submit and wait and fills batch k+1 meanwhile.  Every slot must be waited for before fc_agg_finalize.
 *   meta, reads          fc_batch descriptors and read rows (n_words words per plane)
 *   rn_idx, rn_rows      the N planes as a sparse list: row numbers (ascending or not) and n_words words per listed row, for
 *                        the rows flagged FC_PF_READ_N
 *   q, read_hash         emit: q_a | q_b << 16 and the strand-invariant read hash per row
 *   qname_hash           emit: may be NULL when every row carries fragment fields (descriptor bits 28-31 != 3,3) and a read name
 *                        occurs in one fragment only: the fragment's first row then stands in for the name
 *   idx                  emit: explicit stream position per row, NULL = idx_base + row
 *   out_mode             0 nothing comes back; 1 out_hits = fc_hit[n]; 2 out_hits = int32 (start, end) per row, out_hit_mask /
 *                        out_strand_mask = one bit per row (row i: word i / 32, bit i % 32): has a breakpoint / '-' strand */
typedef struct fc_stream fc_stream;
typedef struct fc_host_batch {
  int64_t n;
  const void* meta;
  const uint32_t* reads;
  const uint32_t* rn_idx;
  const uint32_t* rn_rows;
  int64_t n_rn;
  const uint32_t* q;
  const uint64_t* read_hash;
  const uint64_t* qname_hash;
  const uint64_t* idx;
  uint64_t idx_base;
  int32_t n_words;
  int32_t max_l;
  int32_t emit;
  int32_t out_mode;
  void* out_hits;
  uint32_t* out_hit_mask;
  uint32_t* out_strand_mask;
} fc_host_batch;
int fc_stream_create(fc_ctx* ctx, int32_t n_slots, int64_t cap_rows, int32_t max_words, fc_stream** out);
void fc_stream_destroy(fc_stream* s);
int fc_stream_submit(fc_stream* s, int32_t slot, const fc_scan_params* p, const fc_host_batch* batch);
int fc_stream_wait(fc_stream* s, int32_t slot);
int fc_stream_query(fc_stream* s, int32_t slot); /* 1: the slot is free / its results have landed, 0: in flight, < 0: error */

/* Two-step batches: after fc_batch_host(..., emit = 0) the host inspects the hits and decides which pairs are
 * recorded (find_circ.py:1319-1329: the linear spans of a fragment count only when its back-splices resolved to at most
 * one junction); fc_batch_emit_host records the pairs with h_mask[i] != 0 (NULL = all) from the retained device copy. */
int fc_batch_emit_host(fc_ctx* ctx, const uint8_t* h_mask, uint64_t idx_base);
/* --all-hits (find_circ.py:1312-1317) for the retained batch: h_tie_off = exclusive prefix sum of n_hits (n+1 entries) */
int fc_batch_ties_host(fc_ctx* ctx, const fc_scan_params* p, const int64_t* h_tie_off, fc_hit* h_ties);

/* ------------------------------------------------------------------ junction aggregation
 * fc_agg_emit: turns scan results into fc_jrec records on the device (first tie of every pair with n_hits>0),
 *   appending to the context's record buffer.  Per-pair payload arrays are device pointers.
 * fc_agg_finalize: sort by key + segmented reduce -> fc_junction table (device), returns the count.
 */
int fc_agg_reset(fc_ctx* ctx);
int fc_agg_emit(fc_ctx* ctx, int64_t n, const fc_hit* d_hits, const int32_t* d_chrom, const uint8_t* d_flags,
                const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash,
                const uint64_t* d_qname_hash, const uint8_t* d_mask /* NULL or per-pair 0/1: record this pair */,
                uint64_t idx_base, void* stream);
/* same with an explicit stream position per pair (rows of a batch that are not in stream order) */
int fc_agg_emit_idx(fc_ctx* ctx, int64_t n, const fc_hit* d_hits, const int32_t* d_chrom, const uint8_t* d_flags,
                    const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash,
                    const uint64_t* d_qname_hash, const uint8_t* d_mask, const uint64_t* d_idx, void* stream);
int fc_agg_append(fc_ctx* ctx, int64_t n, const fc_jrec* d_recs, void* stream); /* records built elsewhere (other ranks) */
int fc_agg_append_host(fc_ctx* ctx, int64_t n, const fc_jrec* h_recs);
/* replace the record buffer by n device records (receive side of the exchange); asynchronous, no host synchronisation */
int fc_agg_replace(fc_ctx* ctx, int64_t n, const fc_jrec* d_recs, void* stream);
int64_t fc_agg_n_records(fc_ctx* ctx);
const fc_jrec* fc_agg_records(fc_ctx* ctx); /* device pointer to the record buffer (for the all-to-all) */
/* destination rank of every record: hash(key) % n_ranks (int32 per record) and per-rank counts (int64[n_ranks]) */
int fc_agg_partition(fc_ctx* ctx, int32_t n_ranks, fc_jrec* d_out_sorted_by_rank, int64_t* h_counts, void* stream);
/* Reduce the records per junction (find_circ.py:486-600, 657-690); returns the number of junctions (fc_agg_fetch copies
 * the table to the host, fc_agg_junctions gives the device pointer).  Three implementations with identical results, chosen by
 * the input: sort-free with one global set for the distinct counts (inputs whose set fits L2), sort-free with the distinct
 * counts through partitions + shared-memory sets (larger inputs), sort-based (weight denominators other than 1, 2, 4, 8;
 * positions beyond 2^41).  For tests the environment can force one: FC_AGG_MODE=sort|hash, FC_AGG_SETS=global|part. */
int64_t fc_agg_finalize(fc_ctx* ctx, void* stream);
int fc_agg_fetch(fc_ctx* ctx, int64_t n, fc_junction* h_out); /* sorted by first_idx */
const fc_junction* fc_agg_junctions(fc_ctx* ctx);             /* device pointer, after finalize */

/* ------------------------------------------------------------------ fused emit + exchange over peer memory (one node)
 * The multi-GPU form of SpliceSiteStorage.add (find_circ.py:681-690): junction keys are owned by rank hash(key) % world.
 * Every rank exports its record buffer and counter block (fc_p2p_export -> 128 handle bytes, exchanged by the host with any
 * collective), opens its peers' (fc_p2p_connect) and from then on fc_scan_emit_p2p / fc_agg_emit_p2p write each record
 * directly into the owner's buffer over NVLink: the buffer of rank d is cut into 2 x world slices of capacity/world
 * records, slice (parity, s) receives what source s sends in the steps of that parity, so a source allocates slots with
 * counters in its OWN memory and no atomic crosses the wire.  One step =
 *     fc_agg_reset_async, [fc_scan_emit_p2p | fc_agg_emit_p2p]*, fc_p2p_barrier, fc_agg_finalize
 * with exactly one barrier per step on every rank (it publishes the slice counts and orders the stores before the owners'
 * reduce; the two parities make a barrier before the next step's stores unnecessary).  fc_agg_finalize before the barrier
 * fails with FC_E_STATE; a slice that overflows fails the step with FC_E_NOMEM on the source and on the owner. */
int fc_p2p_export(fc_ctx* ctx, int64_t capacity_records /* per step and rank */, uint8_t* h_handles /* 128 bytes */);
int fc_p2p_connect(fc_ctx* ctx, int32_t world, int32_t rank, const uint8_t* h_all_handles /* world x 128 */,
                   const int64_t* h_capacities /* world */);
/* the same between contexts of ONE process on ONE device (how the single-GPU test suite drives the peer kernels): plain
 * device pointers instead of IPC handles; the barrier of such a context publishes and never waits -- the caller runs the
 * ranks one after the other on one stream (all emits, then all barriers, then the finalizes). */
int fc_p2p_export_local(fc_ctx* ctx, int64_t capacity_records, void** out_recs, void** out_counters);
int fc_p2p_connect_local(fc_ctx* ctx, int32_t world, int32_t rank, void* const* recs, void* const* counters,
                         const int64_t* h_capacities);
int fc_agg_emit_p2p(fc_ctx* ctx, int64_t n, const fc_hit* d_hits, const int32_t* d_chrom, const uint8_t* d_flags,
                    const uint8_t* d_wden, const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash,
                    const uint64_t* d_qname_hash, const uint8_t* d_mask, uint64_t idx_base, void* stream);
/* fc_scan and fc_agg_emit_p2p in one kernel (cf. fc_scan_emit). */
int fc_scan_emit_p2p(fc_ctx* ctx, const fc_scan_params* p, const fc_pairs* pairs, fc_hit* d_out, const uint8_t* d_wden,
                     const int16_t* d_q_a, const int16_t* d_q_b, const uint64_t* d_read_hash, const uint64_t* d_qname_hash,
                     uint64_t idx_base, void* stream);
/* Stream-ordered barrier of all connected ranks over peer memory (one small kernel: every rank publishes its slice counts,
 * bumps an arrival word on every rank and waits for its own).  Every rank must call it exactly once per step.  A rank
 * that has waited longer than the timeout (default 60 s, fc_p2p_set_timeout; the wait is stream-ordered behind the
 * rank's own work, so the timeout must cover the load imbalance between ranks, host ingest included) gives up: the
 * next fc_agg_finalize then fails with FC_E_STATE and the results of that step are invalid. */
int fc_p2p_barrier(fc_ctx* ctx, void* stream);
int fc_p2p_set_timeout(fc_ctx* ctx, double seconds);
int fc_agg_reset_async(fc_ctx* ctx, void* stream);

/* ------------------------------------------------------------------ native SAM ingest (host code)
 * Replaces the pysam record loop, MateSegments, adjacent_segment_pairs and the head of record_hits (find_circ.py:461-469,
 * 976-1140, 1450-1486, 1492-1526, 1560-1574) for fragments of one or two mates whose segments form at most two anchor
 * pairs ("spans") in total -- single-end two-segment reads, mate pairs with one or both mates spliced, a mate with three
 * segments -- and for every fragment without a span (counters only).  Other fragments (three or more spans, a third mate,
 * --no-linear with back-splice and linear spans in one fragment, records the parser cannot interpret) come back as byte
 * ranges for the python implementation of the same logic.  Rows are ready for fc_batch_host_planes; the rows of one
 * fragment are adjacent (back-splice spans first, find_circ.py:1299, 1351) and its fragment record tells the host what the
 * evidence rules of record_hits (find_circ.py:1276-1439) need besides the scan's answers. */
typedef struct fc_ingest fc_ingest;
typedef struct fc_ingest_params {
  int32_t asize, margin, min_uniq_qual, nolinear;
} fc_ingest_params;
/* fragment record flags */
#define FC_FR_UNSPLICED 1u   /* the fragment has an unspliced mate (un_tid / un_pos / un_aend) */
#define FC_FR_OTHER_CHROM 2u /* ... which lies on another chromosome than the first back-splice span's read */
#define FC_FR_BROKEN 4u      /* segments on other chromosomes / strands next to a span that leaves a read end uncovered */
#define FC_FR_TWO_MATES 8u
typedef struct fc_ingest_out {
  int64_t cap;          /* capacity of the per-row and per-fragment arrays (also the stride of the plane arrays) */
  int32_t* chrom;       /* genome chromosome id */
  int32_t* a_start;
  int32_t* b_end;
  int32_t* l;
  uint8_t* flags;
  uint32_t* rlo;        /* [n_words][cap] */
  uint32_t* rhi;
  uint32_t* rn;
  int32_t n_words;
  int32_t max_l;        /* out */
  uint8_t* wden;
  int16_t* q_a;
  int16_t* q_b;
  uint64_t* read_hash;
  uint64_t* qname_hash;
  int64_t* frag_seq;    /* ordinal of the row's fragment in the stream */
  uint8_t* idx_k;       /* the row's place among the fragment's rows: stream position = frag_seq * 64 + idx_k */
  /* one record per fragment that has rows */
  int64_t* f_seq;       /* ordinal of the fragment */
  int32_t* f_row0;      /* its first row */
  uint8_t* f_nsp;       /* spans: 1 or 2, back-splices first */
  uint8_t* f_kind;      /* bit j: span j is a back-splice */
  uint8_t* f_state;     /* bit j: span j has a row (its anchors are unique enough) */
  uint8_t* f_flags;     /* FC_FR_* */
  int32_t* f_un_tid;    /* the unspliced mate: SAM reference index, start, end */
  int32_t* f_un_pos;
  int32_t* f_un_aend;
  int64_t* f_txt_off;   /* [cap][2 mates][name, sequence, qualities]: byte offsets into the parsed text */
  int32_t* f_txt_len;   /* -1: '*', -2: no second mate */
  int64_t cap_complex;
  int64_t* cx_start;    /* byte ranges of fragments left to the python path */
  int64_t* cx_end;
  int64_t* cx_seq;
  int64_t n_rows, n_frag_records, n_complex, n_fragments; /* out */
  double counters[8];   /* out: total_mates, unmapped_reads, unspliced_mates, seg_too_short_skip, circ_junc_not_unique, lin_junc_not_unique */
} fc_ingest_out;
fc_ingest* fc_ingest_create(const fc_ingest_params* p, int32_t n_names, const char* const* names, const int32_t* tid2gid);
void fc_ingest_destroy(fc_ingest* h);
/* a parser that starts inside the stream (one rank of a multi-GPU run takes a byte range of the file, cut where the read
 * name changes): ordinal of its first fragment; at_stream_start = 0 unless its first record is the first of the whole
 * stream -- the only record the reference never checks for the "unmapped" flag (find_circ.py:1462-1463) */
int fc_ingest_set_position(fc_ingest* h, int64_t first_fragment, int32_t at_stream_start);
int64_t fc_ingest_position(fc_ingest* h); /* the ordinal the next fragment will get */
/* parses complete fragments out of `text`; returns the number of bytes consumed (the caller re-submits the rest together
 * with the next chunk; final != 0 flushes the last fragment) or a negative error code */
int64_t fc_ingest_parse(fc_ingest* h, const char* text, int64_t nbytes, int32_t final, fc_ingest_out* out);

/* BAM input (pysam.Samfile(path, 'rb'), find_circ.py:461-469): the file is inflated (BGZF, zlib) and handed out as SAM
 * text lines -- the eleven mandatory columns and the AS / XS tags -- for fc_ingest_parse.  fc_bam_read_text fills `out`
 * (cap >= 64 KiB) with whole lines and returns the bytes written, 0 at the end of the file, FC_E_IO for a broken file. */
typedef struct fc_bam fc_bam;
fc_bam* fc_bam_open(const char* path);
void fc_bam_close(fc_bam* b);
int32_t fc_bam_n_ref(const fc_bam* b);
const char* fc_bam_ref_name(const fc_bam* b, int32_t i);
int64_t fc_bam_ref_length(const fc_bam* b, int32_t i);
int64_t fc_bam_read_text(fc_bam* b, char* out, int64_t cap);

/* The evidence rules of record_hits (find_circ.py:1276-1439, flags of :1319-1344, :1369-1375, :1396-1437) for one batch of
 * the native ingest: the n rows / m fragment records of an fc_ingest_parse call plus the scan's answers for the rows.
 * Every fragment here has at most two spans, so each rule is a comparison between its columns: one host pass in C++.
 * `bit`: the flag bits in this order -- WARN_UNRESOLVED_EXTRA_BACKSPLICE, SUPPORT_CLOSURE, WARN_UNRESOLVED_LINSPLICE,
 * WARN_OUTSIDE_SPLICE_JUNCTION, SUPPORT_INSIDE_SPLICE_JUNCTION, WARN_OTHER_CHROM_MATE, WARN_OUTSIDE_MATE,
 * SUPPORT_INSIDE_MATE, BROKEN_SEGMENTS, WARN_MULTI_BACKSPLICE.  Keys are (chrom id, start, end, minus, kind) rows. */
#define FC_EV_HIT0 1u      /* fc_evidence_out.cls: the first / second span found a breakpoint */
#define FC_EV_HIT1 2u
#define FC_EV_LIN0 4u      /* the first / second span is a linear splice next to the fragment's one back-splice ... */
#define FC_EV_LIN1 8u
#define FC_EV_LIN0_OUT 16u /* ... and lies outside of it */
#define FC_EV_LIN1_OUT 32u
#define FC_EV_UN 64u       /* an unspliced mate next to the one back-splice ... */
#define FC_EV_UN_OUT 128u  /* ... on another chromosome or outside of it */
typedef struct fc_evidence_in {
  int64_t n, m;
  const fc_hit* hits;          /* n */
  const int32_t* chrom;        /* n */
  const uint64_t* qname_hash;  /* n */
  const int64_t* f_seq;        /* m: fc_ingest_out columns of the same names */
  const int32_t* f_row0;
  const uint8_t *f_nsp, *f_kind, *f_state, *f_flags;
  const int32_t *f_un_pos, *f_un_aend;
  const int64_t* f_txt_off;    /* 6 m */
  const int32_t* f_txt_len;    /* 6 m */
  int64_t text_off;            /* added to every text offset (position of the parsed piece in the caller's buffer) */
  int32_t asize;
  uint32_t bit[10];
} fc_evidence_in;
typedef struct fc_evidence_out {
  int64_t counters[4];         /* circ_spliced, circ_no_bp, lin_spliced, lin_no_bp (find_circ.py:1303-1317, 1350-1367) */
  int64_t n_events, n_reads;
  int32_t any_hit;
  uint32_t* W;                 /* m: flag word of the fragment */
  uint8_t* cls;                /* m: FC_EV_* */
  int64_t *key0, *key1, *ck;   /* m x 5: junction of span 0 / span 1 / the (last) back-splice */
  int64_t* ev_key;             /* 2 m x 5: per-junction evidence events (find_circ.py:1325-1327, 1433-1437) ... */
  uint64_t* ev_hash;           /* ... the fragment's name hash ... */
  uint32_t* ev_mask;           /* ... and its flags */
  int64_t* r_seq;              /* 2 m: reads to write (find_circ.py:1439-1447), one per mate: stream position, ... */
  int64_t *r_k0, *r_k1;        /* ... first and (or -1) second junction, */
  int64_t* r_mask;             /* flags, */
  int64_t* r_off3;             /* 2 m x 3: name / sequence / qualities in the text buffer (for fc_text_gather) */
  int32_t* r_len3;
} fc_evidence_out;
int fc_ingest_evidence(const fc_evidence_in* in, fc_evidence_out* out);

/* Distinct rows of an n x width matrix of 64-bit integers in order of first appearance (host code, exact): first[u] = the
 * row where value u appears first, inverse[i] = u for row i; returns the number of distinct rows.  What the writers group by:
 * the (junction, junction, flags) combinations of the reads to print (find_circ.py:1442-1447) and the junctions of the
 * evidence events (:1433-1437). */
int64_t fc_unique_rows(const int64_t* rows, int64_t n, int32_t width, int64_t* first, int32_t* inverse);

/* Spliced reads of the native ingest (write_read, find_circ.py:1442-1447): fc_text_gather copies n x 3 substrings
 * (name, sequence, qualities; off/len row major, len < 0 = absent) of a text buffer back to back into `out` and returns
 * the bytes written; fc_fastq_format turns such a blob into FASTQ records "@<name> <tail>\n<seq>\n+<name> <tail>\n<qual>\n"
 * once the junction names are known (tail = "<junction names> <flags>"; name_idx[i] selects names[name_off[.] ..
 * +name_len[.]]), fills rec_off[0..n] with the offset of every record and returns the bytes written, or the bytes needed
 * when out_cap is too small. */
int64_t fc_text_gather(const char* buf, int64_t n, const int64_t* off, const int32_t* len, char* out);
int64_t fc_fastq_format(const char* blob, int64_t n, const int32_t* len, const int32_t* name_idx, const char* names,
                        const int64_t* name_off, const int32_t* name_len, char* out, int64_t out_cap, int64_t* rec_off);

/* ------------------------------------------------------------------ keyed merge of junction tables
 * The arithmetic of merge_bed.py (merge_bed.py:88-142; column map :117-130): the n rows of ALL input tables (host arrays; row
 * order = file after file, as merge_bed.py reads them) are grouped by (chrom, start, end, strand) on the device; groups are
 * numbered in key order.  Per group: h_support = bit mask of the inputs (h_src, < 64) that have the key (merge_bed.py:80-86),
 * and every numeric column c (h_vals[c * n + row]) reduced over the group's rows in input order with h_op[c] = 0 sum, 1 max,
 * 2 min into h_out[c * n_groups + group].  h_group_of_row tells the host which rows to join for the text columns. */
int fc_merge_tables(fc_ctx* ctx, int64_t n, const uint32_t* h_chrom, const int32_t* h_start, const int32_t* h_end,
                    const uint8_t* h_strand, const uint8_t* h_src, int32_t n_cols, const double* h_vals, const uint8_t* h_op,
                    int64_t* out_n_groups, uint32_t* h_group_of_row, uint64_t* h_support, double* h_out);

/* ------------------------------------------------------------------ utilities */
void* fc_pinned_alloc(int64_t bytes);
void fc_pinned_free(void* p);
int fc_device_sync(fc_ctx* ctx);
/* Declares (after fc_agg_reset*) that every record of this aggregation -- emitted here, appended, or arriving from peer
 * ranks -- has lo <= idx < hi.  With a range about as large as the record count fc_agg_finalize ranks the junctions by
 * discovery order (find_circ.py:684-686) with flags over the range instead of a sort.  A junction whose first record lies
 * outside a declared range is detected: the call then ranks by sort (same result, slower).  Multi-GPU: idx is the position in the whole input stream, the range is the same on
 * every rank. */
int fc_agg_set_idx_range(fc_ctx* ctx, uint64_t lo, uint64_t hi);
/* Device time of the stages of the last fc_agg_finalize() call on its sort-free path, measured with CUDA events on the
 * caller's stream while switched on: out_us[0..5] = table clears, accumulate kernel, distinct-count kernel over the
 * partitions (0 for inputs whose set fits L2), first-record marks, finish kernel, counter copy (bench.py's roofline for the aggregation kernel; no counterpart in the reference). */
int fc_agg_set_timing(fc_ctx* ctx, int32_t on);
int fc_agg_get_timing(fc_ctx* ctx, float* out_us);
/* number of kernels this library has launched in this context (bench.py's gpu_launches) */
int64_t fc_launch_count(fc_ctx* ctx);
/* the hash functions the host must use for fc_jrec.read_hash / qname_hash (FNV-1a 64 + finaliser) */
uint64_t fc_hash_bytes(const uint8_t* p, int64_t n);
/* strand-invariant read hash + palindrome flag: min(hash(seq.upper()), hash(revcomp)) */
uint64_t fc_hash_read(const uint8_t* seq, int64_t n, int32_t* is_palindrome);
/* vectorised: rows of a fixed-stride matrix */
int fc_hash_reads_host(int64_t n, const uint8_t* h_seq, int32_t stride, const int32_t* h_len, uint64_t* h_out,
                       uint8_t* h_pal);

/* the same on device rows (one thread per read); d_len may be NULL: every row holds fixed_len letters */
int fc_hash_reads_device(fc_ctx* ctx, int64_t n, const uint8_t* d_seq, int32_t stride, const int32_t* d_len, int32_t fixed_len,
                         uint64_t* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FINDCIRC_B200_H */
