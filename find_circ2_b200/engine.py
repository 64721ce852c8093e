"""
Engine -- python face of one libfindcirc_b200 context (one GPU).

Mirrors the three objects of the reference that sit on the hot path:
  genome store      Track(options.genome, accessor=GenomeAccessor)        find_circ.py:435-436
  breakpoint scan   JunctionSpan.find_breakpoints()                       find_circ.py:854-974
  junction tables   SpliceSiteStorage("circ") / SpliceSiteStorage("lin")  find_circ.py:1142-1143
All arithmetic happens in CUDA kernels behind the C ABI (include/findcirc_b200.h); torch is used only to hold
device buffers and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import HIT_DTYPE, JREC_DTYPE, JUNCTION_DTYPE, Batch, FindCircError, HostBatch, Pairs, ScanParams, ptr

SIG_LETTERS = "ACGTN"


def decode_signal(code: int) -> str:
    return "".join(SIG_LETTERS[(code >> (3 * k)) & 7] for k in range(4))


class HostStream:
    """fc_stream: batches in fc_batch layout go from (pinned) host arrays through copy, scan, record and back without the host
    waiting in between; the caller fills batch k+1 while batch k is in flight (find_circ.py:1535-1574 as a pipeline)"""

    def __init__(self, eng: "Engine", n_slots: int, cap_rows: int, max_words: int):
        self.eng = eng
        self.n_slots = n_slots
        h = C.c_void_p()
        eng._check(eng.lib.fc_stream_create(eng.h, n_slots, int(cap_rows), int(max_words), C.byref(h)))
        self.h = h
        self._keep = [None] * n_slots  # the arrays of a batch stay alive while it is in flight

    def submit(self, slot: int, n: int, meta, reads, n_words: int, max_l: int, q=None, read_hash=None, qname_hash=None, idx=None,
               idx_base: int = 0, rn_idx=None, rn_rows=None, emit: bool = True, out_mode: int = 0, out_hits=None, out_hit_mask=None,
               out_strand_mask=None):
        hb = HostBatch(int(n), ptr(meta), ptr(reads), ptr(rn_idx), ptr(rn_rows), 0 if rn_idx is None else len(rn_idx), ptr(q),
                       ptr(read_hash), ptr(qname_hash), ptr(idx), int(idx_base), int(n_words), int(max_l), int(bool(emit)), int(out_mode),
                       ptr(out_hits), ptr(out_hit_mask), ptr(out_strand_mask))
        self._keep[slot] = (meta, reads, q, read_hash, qname_hash, idx, rn_idx, rn_rows, out_hits, out_hit_mask, out_strand_mask)
        self.eng._check(self.eng.lib.fc_stream_submit(self.h, slot, C.byref(self.eng.params), C.byref(hb)))

    def wait(self, slot: int):
        self.eng._check(self.eng.lib.fc_stream_wait(self.h, slot))
        self._keep[slot] = None

    def wait_all(self):
        for k in range(self.n_slots):
            self.wait(k)

    def query(self, slot: int) -> bool:
        return bool(self.eng._check(self.eng.lib.fc_stream_query(self.h, slot)))

    def close(self):
        if getattr(self, "h", None):
            self.eng.lib.fc_stream_destroy(self.h)
            self.h = None


class Engine:
    def __init__(self, device: int = 0, asize: int = 15, margin: int = 2, maxdist: int = 2, noncanonical: bool = False,
                 strandpref: bool = False):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.fc_ctx_create(device, C.byref(h))
        if rc != 0:
            raise FindCircError(rc, self.lib.fc_last_error(None).decode())
        self.h = h
        self.device = device
        self.params = ScanParams(asize, margin, maxdist, int(noncanonical), int(strandpref))
        self.chrom_names: List[str] = []
        self.chrom_sizes: List[int] = []
        self._chrom_ids = {}
        self._fetch_ptr, self._fetch_bytes = None, 0  # pinned landing buffer of agg_fetch

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc is not None and rc < 0:
            raise FindCircError(rc, self.lib.fc_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "_fetch_ptr", None):
            self.lib.fc_pinned_free(self._fetch_ptr)
            self._fetch_ptr, self._fetch_bytes = None, 0
        if getattr(self, "h", None):
            self.lib.fc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def eff(self) -> int:
        return self.params.asize - self.params.margin

    def launch_count(self) -> int:
        return int(self.lib.fc_launch_count(self.h))

    def sync(self):
        self._check(self.lib.fc_device_sync(self.h))

    # ------------------------------------------------------------------ genome store
    def _refresh_chroms(self):
        n = self.lib.fc_genome_n_chrom(self.h)
        buf = C.create_string_buffer(4096)
        self.chrom_names, self.chrom_sizes = [], []
        for i in range(n):
            self._check(self.lib.fc_genome_chrom_name(self.h, i, buf, 4096))
            self.chrom_names.append(buf.value.decode())
            self.chrom_sizes.append(int(self.lib.fc_genome_chrom_size(self.h, i)))
        self._chrom_ids = {n: i for i, n in enumerate(self.chrom_names)}

    def load_genome_fasta(self, path: str):
        self._check(self.lib.fc_genome_load_fasta(self.h, path.encode()))
        self._refresh_chroms()

    def load_genome_arrays(self, names: Sequence[str], seqs: Sequence[np.ndarray]):
        """chromosomes as uint8 ASCII arrays (synthetic genomes)"""
        seqs = [np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
        n = len(seqs)
        c_names = (C.c_char_p * n)(*[s.encode() for s in names])
        c_seqs = (C.c_void_p * n)(*[s.ctypes.data for s in seqs])
        sizes = np.array([len(s) for s in seqs], dtype=np.int64)
        self._check(self.lib.fc_genome_load_ascii(self.h, n, c_names, c_seqs, sizes.ctypes.data))
        self._refresh_chroms()

    def share_genome(self, other: "Engine"):
        """use the device store of another engine of this process on the same device (no second copy)"""
        self._check(self.lib.fc_genome_share(self.h, other.h))
        self._refresh_chroms()

    def chrom_id(self, name: str) -> int:
        """KeyError for unknown chromosomes, as the reference raises (find_circ.py:193)"""
        return self._chrom_ids[name]

    def genome_stats(self):
        s = np.zeros(4, dtype=np.int64)
        self._check(self.lib.fc_genome_stats(self.h, s.ctypes.data))
        return {"bases": int(s[0]), "n": int(s[1]), "other_as_n": int(s[2]), "device_bytes": int(s[3])}

    def fetch(self, chrom: int, start: int, end: int) -> str:
        """genome.get(chrom,start,end,'+').upper() decoded from the device store"""
        n = max(0, end - start)
        buf = np.zeros(n, dtype=np.uint8)
        self._check(self.lib.fc_genome_fetch(self.h, chrom, start, end, buf.ctypes.data))
        return buf.tobytes().decode()

    # ------------------------------------------------------------------ scan, host buffers in / out
    def scan_host(self, chrom, a_start, b_end, l, flags, internal, out: Optional[np.ndarray] = None) -> np.ndarray:
        """One batch through H2D -> pack -> scan -> D2H.  `internal` is an [n, stride] uint8 ASCII matrix."""
        n = len(chrom)
        if out is None:
            out = np.zeros(n, dtype=HIT_DTYPE)
        if n == 0:
            return out
        internal = np.ascontiguousarray(internal, dtype=np.uint8)
        if internal.ndim != 2 or internal.shape[0] != n:
            raise ValueError("internal must be [n, stride]")
        stride = max(int(internal.shape[1]), 1)
        if internal.shape[1] == 0:
            internal = np.zeros((n, 1), dtype=np.uint8)
        args = [np.ascontiguousarray(x, dtype=t) for x, t in
                ((chrom, np.int32), (a_start, np.int32), (b_end, np.int32), (l, np.int32), (flags, np.uint8))]
        self._check(self.lib.fc_scan_host(self.h, C.byref(self.params), n, *[a.ctypes.data for a in args],
                                          internal.ctypes.data, stride, out.ctypes.data))
        return out

    def batch_host(self, chrom, a_start, b_end, l, flags, internal, wden, q_a, q_b, read_hash, qname_hash, idx_base,
                   emit=True, out: Optional[np.ndarray] = None, want_hits=True, idx=None) -> Optional[np.ndarray]:
        """scan one batch and append its junction records to the device aggregator (host buffers in, hits out)"""
        n = len(chrom)
        if want_hits and out is None:
            out = np.zeros(n, dtype=HIT_DTYPE)
        if n == 0:
            return out
        internal = np.ascontiguousarray(internal, dtype=np.uint8)
        if internal.shape[1] == 0:
            internal = np.zeros((n, 1), dtype=np.uint8)
        stride = int(internal.shape[1])
        a = [np.ascontiguousarray(x, dtype=t) for x, t in (
            (chrom, np.int32), (a_start, np.int32), (b_end, np.int32), (l, np.int32), (flags, np.uint8))]
        pay = [np.ascontiguousarray(x, dtype=t) for x, t in (
            (wden, np.uint8), (q_a, np.int16), (q_b, np.int16), (read_hash, np.uint64), (qname_hash, np.uint64))]
        h_idx = None if idx is None else np.ascontiguousarray(idx, dtype=np.uint64)
        self._check(self.lib.fc_batch_host_idx(
            self.h, C.byref(self.params), n, *[x.ctypes.data for x in a], internal.ctypes.data, stride,
            *[x.ctypes.data for x in pay], None if h_idx is None else h_idx.ctypes.data, int(idx_base), int(bool(emit)),
            out.ctypes.data if want_hits else None))
        return out

    def batch_host_planes(self, n, chrom, a_start, b_end, l, flags, rlo, rhi, rn, n_words, plane_stride, max_l, wden, q_a, q_b,
                          read_hash, qname_hash, idx=None, idx_base=0, emit=True, out: Optional[np.ndarray] = None):
        """rows whose internal read part is already packed as bit planes (native ingest); arrays may be longer than n"""
        if out is None:
            out = np.zeros(n, dtype=HIT_DTYPE)
        if n == 0:
            return out
        self._check(self.lib.fc_batch_host_planes(
            self.h, C.byref(self.params), n, chrom.ctypes.data, a_start.ctypes.data, b_end.ctypes.data, l.ctypes.data,
            flags.ctypes.data, rlo.ctypes.data, rhi.ctypes.data, rn.ctypes.data, int(n_words), int(plane_stride), int(max_l),
            wden.ctypes.data, q_a.ctypes.data, q_b.ctypes.data, read_hash.ctypes.data, qname_hash.ctypes.data,
            None if idx is None else idx.ctypes.data, int(idx_base), int(bool(emit)), out.ctypes.data))
        return out

    # ------------------------------------------------------------------ scan, device buffers (torch tensors)
    def pack_reads(self, d_ascii, stride: int, d_l, n_words: int, d_planes, d_flags, stream=0):
        """d_planes: int32 device tensor of 3*n_words*n words (lo plane, hi plane, N plane)"""
        n = d_l.numel()
        lo, hi, nn = self.plane_ptrs(d_planes, n, n_words)
        self._check(self.lib.fc_pack_reads(self.h, n, ptr(d_ascii), stride, ptr(d_l), n_words, lo, hi, nn,
                                           ptr(d_flags), stream))

    @staticmethod
    def n_words_for(max_l: int) -> int:
        return max(1, (max_l + 31) // 32)

    @staticmethod
    def plane_ptrs(d_planes, n, n_words):
        base = ptr(d_planes)
        step = 4 * n * n_words
        return base, base + step, base + 2 * step

    def make_pairs(self, n, d_chrom, d_a_start, d_b_end, d_l, d_flags, d_planes, n_words, max_l) -> Pairs:
        lo, hi, nn = self.plane_ptrs(d_planes, n, n_words)
        return Pairs(n, ptr(d_chrom), ptr(d_a_start), ptr(d_b_end), ptr(d_l), ptr(d_flags), lo, hi, nn, n_words, max_l, 0)

    def scan(self, pairs: Pairs, d_out, stream=0):
        self._check(self.lib.fc_scan(self.h, C.byref(self.params), C.byref(pairs), ptr(d_out), stream))

    def scan_emit(self, pairs: Pairs, d_out, d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, stream=0, d_idx=None):
        """scan + record the pairs that found a breakpoint, in one kernel (fc_scan_emit)"""
        self._check(self.lib.fc_scan_emit(self.h, C.byref(self.params), C.byref(pairs), ptr(d_out), ptr(d_wden), ptr(d_q_a),
                                          ptr(d_q_b), ptr(d_read_hash), ptr(d_qname_hash), idx_base, ptr(d_idx), stream))

    def scan_emit_p2p(self, pairs: Pairs, d_out, d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, stream=0):
        """scan + record into the owner ranks' buffers over peer memory, in one kernel (fc_scan_emit_p2p)"""
        self._check(self.lib.fc_scan_emit_p2p(self.h, C.byref(self.params), C.byref(pairs), ptr(d_out), ptr(d_wden), ptr(d_q_a),
                                              ptr(d_q_b), ptr(d_read_hash), ptr(d_qname_hash), idx_base, stream))

    # ---- packed batches (fc_batch): the layout the scan kernels read
    def batch_words(self, max_l: int) -> int:
        return int(self.lib.fc_batch_words(int(max_l)))

    def pack_batch(self, pairs: Pairs, d_meta, d_reads, d_rn, d_wden=None, d_q_a=None, d_q_b=None, d_q=None, d_frag=None, stream=0) -> Batch:
        """convert an fc_pairs batch on the device (fc_batch_pack); d_meta: 16 bytes per pair, d_reads: 2 * batch_words words
        per pair, d_rn: batch_words words per pair, d_q (with d_q_a / d_q_b): one word per pair"""
        self._check(self.lib.fc_batch_pack(self.h, C.byref(pairs), ptr(d_wden), ptr(d_q_a), ptr(d_q_b), ptr(d_frag), ptr(d_meta),
                                           ptr(d_reads), ptr(d_rn), ptr(d_q), stream))
        return Batch(pairs.n, ptr(d_meta), ptr(d_reads), ptr(d_rn), self.batch_words(pairs.max_l), pairs.max_l)

    def scan_batch(self, batch: Batch, d_out, stream=0):
        self._check(self.lib.fc_scan_batch(self.h, C.byref(self.params), C.byref(batch), ptr(d_out), stream))

    def scan_emit_batch(self, batch: Batch, d_out, d_q, d_read_hash, d_qname_hash, idx_base, stream=0, d_idx=None):
        """scan + record in one kernel; on a context connected to peers the records go to the ranks that own their keys"""
        self._check(self.lib.fc_scan_emit_batch(self.h, C.byref(self.params), C.byref(batch), ptr(d_out), ptr(d_q), ptr(d_read_hash),
                                                ptr(d_qname_hash), idx_base, ptr(d_idx), stream))

    def chrom_offset(self, i: int) -> int:
        """genome coordinate of base 0 of chromosome i (what fc_batch descriptors are made of)"""
        return int(self.lib.fc_genome_chrom_offset(self.h, int(i)))

    def scan_ties(self, pairs: Pairs, d_hits, d_tie_off, d_ties, stream=0):
        self._check(self.lib.fc_scan_ties(self.h, C.byref(self.params), C.byref(pairs), ptr(d_hits), ptr(d_tie_off),
                                          ptr(d_ties), stream))

    # ------------------------------------------------------------------ aggregation
    def agg_reset(self):
        self._check(self.lib.fc_agg_reset(self.h))

    def agg_emit(self, n, d_hits, d_chrom, d_flags, d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, stream=0,
                 d_mask=None):
        self._check(self.lib.fc_agg_emit(self.h, n, ptr(d_hits), ptr(d_chrom), ptr(d_flags), ptr(d_wden), ptr(d_q_a),
                                         ptr(d_q_b), ptr(d_read_hash), ptr(d_qname_hash), ptr(d_mask), idx_base, stream))

    def batch_emit(self, mask, idx_base):
        """record the pairs of the last batch_host(emit=False) call whose mask byte is non-zero (None = all)"""
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._check(self.lib.fc_batch_emit_host(self.h, None if m is None else m.ctypes.data, int(idx_base)))

    def batch_ties(self, n_hits: np.ndarray) -> (np.ndarray, np.ndarray):
        """every tie of every pair of the last batch (--all-hits); returns (offsets[n+1], ties)"""
        off = np.zeros(len(n_hits) + 1, dtype=np.int64)
        off[1:] = np.cumsum(n_hits.astype(np.int64))
        ties = np.zeros(int(off[-1]), dtype=HIT_DTYPE)
        if len(ties):
            self._check(self.lib.fc_batch_ties_host(self.h, C.byref(self.params), off.ctypes.data, ties.ctypes.data))
        return off, ties

    def agg_append_host(self, recs: np.ndarray):
        recs = np.ascontiguousarray(recs, dtype=JREC_DTYPE)
        self._check(self.lib.fc_agg_append_host(self.h, len(recs), recs.ctypes.data))

    def agg_append_device(self, n, d_recs, stream=0):
        self._check(self.lib.fc_agg_append(self.h, n, ptr(d_recs), stream))

    # ---- fused emit + exchange over peer memory (multi-GPU) ----
    def p2p_export(self, capacity_records: int) -> np.ndarray:
        h = np.zeros(128, dtype=np.uint8)
        self._check(self.lib.fc_p2p_export(self.h, int(capacity_records), h.ctypes.data))
        return h

    def p2p_connect(self, world: int, rank: int, handles: np.ndarray, capacities: np.ndarray):
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        capacities = np.ascontiguousarray(capacities, dtype=np.int64)
        self._check(self.lib.fc_p2p_connect(self.h, world, rank, handles.ctypes.data, capacities.ctypes.data))

    def p2p_export_local(self, capacity_records: int):
        """(record buffer, counter block) device pointers for contexts of the same process (fc_p2p_export_local)"""
        recs, cnt = C.c_void_p(), C.c_void_p()
        self._check(self.lib.fc_p2p_export_local(self.h, int(capacity_records), C.byref(recs), C.byref(cnt)))
        return recs.value, cnt.value

    def p2p_connect_local(self, world: int, rank: int, recs, counters, capacities):
        r = (C.c_void_p * world)(*recs)
        c = (C.c_void_p * world)(*counters)
        capacities = np.ascontiguousarray(capacities, dtype=np.int64)
        self._check(self.lib.fc_p2p_connect_local(self.h, world, rank, r, c, capacities.ctypes.data))

    def p2p_set_timeout(self, seconds: float):
        self._check(self.lib.fc_p2p_set_timeout(self.h, float(seconds)))

    def agg_emit_p2p(self, n, d_hits, d_chrom, d_flags, d_wden, d_q_a, d_q_b, d_read_hash, d_qname_hash, idx_base, stream=0,
                     d_mask=None):
        self._check(self.lib.fc_agg_emit_p2p(self.h, n, ptr(d_hits), ptr(d_chrom), ptr(d_flags), ptr(d_wden), ptr(d_q_a),
                                             ptr(d_q_b), ptr(d_read_hash), ptr(d_qname_hash), ptr(d_mask), idx_base, stream))

    def p2p_barrier(self, stream=0):
        """stream-ordered barrier of the connected ranks over peer memory (fc_p2p_barrier)"""
        self._check(self.lib.fc_p2p_barrier(self.h, stream))

    def agg_reset_async(self, stream=0):
        self._check(self.lib.fc_agg_reset_async(self.h, stream))

    def agg_replace_device(self, n, d_recs, stream=0):
        self._check(self.lib.fc_agg_replace(self.h, n, ptr(d_recs), stream))

    def agg_n_records(self) -> int:
        return int(self._check(self.lib.fc_agg_n_records(self.h)))

    def agg_records_ptr(self) -> int:
        return int(self.lib.fc_agg_records(self.h) or 0)

    def agg_partition(self, n_ranks: int, d_out, stream=0) -> np.ndarray:
        counts = np.zeros(n_ranks, dtype=np.int64)
        self._check(self.lib.fc_agg_partition(self.h, n_ranks, ptr(d_out), counts.ctypes.data, stream))
        return counts

    def agg_set_idx_range(self, lo: int, hi: int):
        """every record of this aggregation (local, appended, from peers) has lo <= idx < hi (call after agg_reset*)"""
        self._check(self.lib.fc_agg_set_idx_range(self.h, int(lo), int(hi)))

    def agg_set_timing(self, on: bool):
        self._check(self.lib.fc_agg_set_timing(self.h, 1 if on else 0))

    def agg_get_timing(self) -> dict:
        """device time (us) of the stages of the last agg_finalize on its sort-free path"""
        out = np.zeros(6, dtype=np.float32)
        self._check(self.lib.fc_agg_get_timing(self.h, out.ctypes.data))
        return dict(zip(("clear", "accumulate", "distinct", "mark", "finish", "copy"), (float(v) for v in out)))

    def agg_finalize(self, stream=0) -> int:
        return int(self._check(self.lib.fc_agg_finalize(self.h, stream)))

    def agg_fetch(self, n: int, copy: bool = True) -> np.ndarray:
        """the junction table of the last agg_finalize; lands in a pinned host buffer of the engine (a pageable
        destination costs a staged copy).  copy=False returns a view that the next agg_fetch overwrites"""
        if n == 0:
            return np.zeros(0, dtype=JUNCTION_DTYPE)
        need = n * JUNCTION_DTYPE.itemsize
        if need > self._fetch_bytes:
            if self._fetch_ptr:
                self.lib.fc_pinned_free(self._fetch_ptr)
            self._fetch_bytes = max(need + need // 2, 1 << 20)
            self._fetch_ptr = self.lib.fc_pinned_alloc(self._fetch_bytes)
            if not self._fetch_ptr:
                self._fetch_bytes = 0
                raise MemoryError("pinned host allocation of %d bytes failed" % need)
        self._check(self.lib.fc_agg_fetch(self.h, n, self._fetch_ptr))
        view = np.frombuffer((C.c_char * need).from_address(self._fetch_ptr), dtype=JUNCTION_DTYPE, count=n)
        return view.copy() if copy else view

    # ------------------------------------------------------------------ hashing (host helpers of the ABI)
    def hash_reads(self, seqs: np.ndarray, lens: np.ndarray) -> np.ndarray:
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        out = np.zeros(len(lens), dtype=np.uint64)
        self._check(self.lib.fc_hash_reads_host(len(lens), seqs.ctypes.data, seqs.shape[1], lens.ctypes.data,
                                                out.ctypes.data, None))
        return out

    def hash_reads_device(self, d_seqs, fixed_len: int, d_out, stream=0):
        """strand-invariant read hashes of device rows ([n, stride] uint8 tensor -> int64/uint64 tensor of n)"""
        n, stride = int(d_seqs.shape[0]), int(d_seqs.shape[1])
        self._check(self.lib.fc_hash_reads_device(self.h, n, ptr(d_seqs), stride, None, int(fixed_len), ptr(d_out), stream))

    def hash_read(self, seq: bytes) -> int:
        return int(self.lib.fc_hash_read(seq, len(seq), None))

    def hash_bytes(self, b: bytes) -> int:
        return int(self.lib.fc_hash_bytes(b, len(b)))
