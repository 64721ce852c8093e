"""
ctypes face of the native SAM ingest (csrc/ingest.cu, fc_ingest_*): parses chunks of SAM text into struct-of-arrays rows
plus one record per fragment, for fragments of one or two mates with at most two anchor pairs in total; everything else
comes back as byte ranges for the python path (pipeline.Run.add_fragment).  Host code only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class IngestParams(C.Structure):
    _fields_ = [("asize", C.c_int32), ("margin", C.c_int32), ("min_uniq_qual", C.c_int32), ("nolinear", C.c_int32)]


_P = C.c_void_p


class IngestOut(C.Structure):
    _fields_ = [
        ("cap", C.c_int64), ("chrom", _P), ("a_start", _P), ("b_end", _P), ("l", _P), ("flags", _P), ("rlo", _P), ("rhi", _P),
        ("rn", _P), ("n_words", C.c_int32), ("max_l", C.c_int32), ("wden", _P), ("q_a", _P), ("q_b", _P), ("read_hash", _P),
        ("qname_hash", _P), ("frag_seq", _P), ("idx_k", _P),
        ("f_seq", _P), ("f_row0", _P), ("f_nsp", _P), ("f_kind", _P), ("f_state", _P), ("f_flags", _P), ("f_un_tid", _P),
        ("f_un_pos", _P), ("f_un_aend", _P), ("f_txt_off", _P), ("f_txt_len", _P),
        ("cap_complex", C.c_int64), ("cx_start", _P), ("cx_end", _P), ("cx_seq", _P),
        ("n_rows", C.c_int64), ("n_frag_records", C.c_int64), ("n_complex", C.c_int64), ("n_fragments", C.c_int64),
        ("counters", C.c_double * 8),
    ]


FR_UNSPLICED, FR_OTHER_CHROM, FR_BROKEN, FR_TWO_MATES = 1, 2, 4, 8  # fc_ingest_out.f_flags

ROW_FIELDS = (("chrom", np.int32), ("a_start", np.int32), ("b_end", np.int32), ("l", np.int32), ("flags", np.uint8),
              ("wden", np.uint8), ("q_a", np.int16), ("q_b", np.int16), ("read_hash", np.uint64), ("qname_hash", np.uint64),
              ("frag_seq", np.int64), ("idx_k", np.uint8))
FRAG_FIELDS = (("f_seq", np.int64), ("f_row0", np.int32), ("f_nsp", np.uint8), ("f_kind", np.uint8), ("f_state", np.uint8),
               ("f_flags", np.uint8), ("f_un_tid", np.int32), ("f_un_pos", np.int32), ("f_un_aend", np.int32))

COUNTER_NAMES = ("total_mates", "unmapped_reads", "unspliced_mates", "seg_too_short_skip", "circ_junc_not_unique",
                 "lin_junc_not_unique")


class EvidenceIn(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("hits", _P), ("chrom", _P), ("qname_hash", _P), ("f_seq", _P), ("f_row0", _P),
                ("f_nsp", _P), ("f_kind", _P), ("f_state", _P), ("f_flags", _P), ("f_un_pos", _P), ("f_un_aend", _P),
                ("f_txt_off", _P), ("f_txt_len", _P), ("text_off", C.c_int64), ("asize", C.c_int32), ("bit", C.c_uint32 * 10)]


class EvidenceOut(C.Structure):
    _fields_ = [("counters", C.c_int64 * 4), ("n_events", C.c_int64), ("n_reads", C.c_int64), ("any_hit", C.c_int32), ("W", _P),
                ("cls", _P), ("key0", _P), ("key1", _P), ("ck", _P), ("ev_key", _P), ("ev_hash", _P), ("ev_mask", _P),
                ("r_seq", _P), ("r_k0", _P), ("r_k1", _P), ("r_mask", _P), ("r_off3", _P), ("r_len3", _P)]


EV_HIT0, EV_HIT1, EV_LIN0, EV_LIN1, EV_LIN0_OUT, EV_LIN1_OUT, EV_UN, EV_UN_OUT = 1, 2, 4, 8, 16, 32, 64, 128  # fc_evidence_out.cls
# the order fc_ingest_evidence expects the flag bits in
EVIDENCE_FLAGS = ("WARN_UNRESOLVED_EXTRA_BACKSPLICE", "SUPPORT_CLOSURE", "WARN_UNRESOLVED_LINSPLICE", "WARN_OUTSIDE_SPLICE_JUNCTION",
                  "SUPPORT_INSIDE_SPLICE_JUNCTION", "WARN_OTHER_CHROM_MATE", "WARN_OUTSIDE_MATE", "SUPPORT_INSIDE_MATE",
                  "BROKEN_SEGMENTS", "WARN_MULTI_BACKSPLICE")


def evidence(lib, a: dict, hits: np.ndarray, n: int, m: int, text_off: int, asize: int, flag_bit: dict) -> dict:
    """fc_ingest_evidence on the arrays of one fc_ingest_parse call (`a`) and the scan's hits for its n rows: counters,
    per-fragment flag words / classes / junction keys, the evidence events and the reads to write (see the header)"""
    i = EvidenceIn()
    i.n, i.m, i.text_off, i.asize = n, m, text_off, asize
    hits = np.ascontiguousarray(hits)
    i.hits = hits.ctypes.data
    for k in ("chrom", "qname_hash", "f_seq", "f_row0", "f_nsp", "f_kind", "f_state", "f_flags", "f_un_pos", "f_un_aend", "f_txt_off",
              "f_txt_len"):
        setattr(i, k, a[k].ctypes.data)
    for k, name in enumerate(EVIDENCE_FLAGS):
        i.bit[k] = flag_bit[name]
    r = {"W": np.empty(m, np.uint32), "cls": np.empty(m, np.uint8), "key0": np.empty((m, 5), np.int64), "key1": np.empty((m, 5), np.int64),
         "ck": np.empty((m, 5), np.int64), "ev_key": np.empty((2 * m, 5), np.int64), "ev_hash": np.empty(2 * m, np.uint64),
         "ev_mask": np.empty(2 * m, np.uint32), "r_seq": np.empty(2 * m, np.int64), "r_k0": np.empty((2 * m, 5), np.int64),
         "r_k1": np.empty((2 * m, 5), np.int64), "r_mask": np.empty(2 * m, np.int64), "r_off3": np.empty((2 * m, 3), np.int64),
         "r_len3": np.empty((2 * m, 3), np.int32)}
    o = EvidenceOut()
    for k, arr in r.items():
        setattr(o, k, arr.ctypes.data)
    rc = lib.fc_ingest_evidence(C.byref(i), C.byref(o))
    if rc != 0:
        raise RuntimeError("fc_ingest_evidence failed (%d)" % rc)
    ne, nr = int(o.n_events), int(o.n_reads)
    # (events are few: copies, so that the full-size buffers do not stay alive behind small views)
    for k in ("ev_key", "ev_hash", "ev_mask"):
        r[k] = r[k][:ne].copy()
    for k in ("r_seq", "r_k0", "r_k1", "r_mask", "r_off3", "r_len3"):
        r[k] = r[k][:nr] if 2 * nr >= len(r[k]) else r[k][:nr].copy()
    r["counters"] = [int(o.counters[k]) for k in range(4)]
    r["any_hit"] = bool(o.any_hit)
    return r


def text_address(buf) -> int:
    """address of the first byte of a bytes / bytearray object (which the caller keeps alive and does not resize)"""
    if isinstance(buf, bytearray):
        return C.addressof(C.c_char.from_buffer(buf))
    return C.cast(C.c_char_p(buf), C.c_void_p).value


class NativeIngest(object):
    def __init__(self, asize, margin, min_uniq_qual, nolinear, names, tid2gid, cap=1 << 18, n_words=8, cap_complex=1 << 16,
                 first_fragment=0, at_stream_start=True):
        self.lib = _lib.load()
        p = IngestParams(asize, margin, min_uniq_qual, int(bool(nolinear)))
        c_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
        t2g = np.ascontiguousarray(tid2gid, dtype=np.int32)
        self.h = self.lib.fc_ingest_create(C.byref(p), len(names), c_names, t2g.ctypes.data)
        if not self.h:
            raise RuntimeError("fc_ingest_create failed")
        if first_fragment or not at_stream_start:
            if self.lib.fc_ingest_set_position(self.h, int(first_fragment), int(bool(at_stream_start))) != 0:
                raise RuntimeError("fc_ingest_set_position failed")
        self.cap, self.n_words = cap, n_words
        a = self.a = {}
        for name, dt in ROW_FIELDS + FRAG_FIELDS:
            a[name] = np.zeros(cap, dtype=dt)
        a["f_txt_off"] = np.zeros(6 * cap, dtype=np.int64)
        a["f_txt_len"] = np.zeros(6 * cap, dtype=np.int32)
        for name in ("rlo", "rhi", "rn"):
            a[name] = np.zeros(n_words * cap, dtype=np.uint32)
        for name in ("cx_start", "cx_end", "cx_seq"):
            a[name] = np.zeros(cap_complex, dtype=np.int64)
        o = self.out = IngestOut()
        o.cap, o.n_words, o.cap_complex = cap, n_words, cap_complex
        for name, arr in a.items():
            setattr(o, name, arr.ctypes.data)

    def set_position(self, first_fragment: int, at_stream_start: bool):
        if self.lib.fc_ingest_set_position(self.h, int(first_fragment), int(bool(at_stream_start))) != 0:
            raise RuntimeError("fc_ingest_set_position failed")

    def next_fragment(self) -> int:
        return int(self.lib.fc_ingest_position(self.h))

    def snapshot(self, n: int, m: int) -> dict:
        """copies of the first n rows / m fragment records of the output arrays (planes as [n_words][n]): what a worker
        thread hands over"""
        out = {}
        for k, _ in ROW_FIELDS:
            out[k] = self.a[k][:n].copy()
        for k, _ in FRAG_FIELDS:
            out[k] = self.a[k][:m].copy()
        out["f_txt_off"] = self.a["f_txt_off"][:6 * m].copy()
        out["f_txt_len"] = self.a["f_txt_len"][:6 * m].copy()
        for k in ("rlo", "rhi", "rn"):
            out[k] = np.ascontiguousarray(self.a[k].reshape(self.n_words, self.cap)[:, :n]).reshape(-1)
        return out

    def parse(self, buf: bytes, offset: int, final: bool, end: int = -1) -> int:
        """parse buf[offset:]; returns bytes consumed.  Results are in self.out / self.a until the next call."""
        base = text_address(buf)
        n = self.lib.fc_ingest_parse(self.h, base + offset, (len(buf) if end < 0 else end) - offset, int(final), C.byref(self.out))
        if n < 0:
            raise RuntimeError("fc_ingest_parse failed (%d)" % n)
        return int(n)

    def close(self):
        if self.h:
            self.lib.fc_ingest_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BamText(object):
    """a BAM file as a stream of SAM text records (csrc/bam.cu): `.names` / `.lengths` of the header, `.read(n)` like a
    binary file -- whole lines only, b"" at the end.  A reader thread inflates the next chunk while the caller parses the
    current one (both are C calls that release the interpreter lock)."""

    def __init__(self, path: str, prefetch: int = 2):
        self.lib = _lib.load()
        self.h = self.lib.fc_bam_open(path.encode())
        if not self.h:
            raise IOError("cannot read '%s' as BAM" % path)
        n = self.lib.fc_bam_n_ref(self.h)
        self.names = [self.lib.fc_bam_ref_name(self.h, i).decode("latin-1") for i in range(n)]
        self.lengths = [int(self.lib.fc_bam_ref_length(self.h, i)) for i in range(n)]
        self._prefetch = prefetch
        self._queue = None
        self._thread = None
        self._stop = False
        self._done = False

    def _fill(self, n: int) -> bytes:
        """the next chunk of at most n bytes (synchronously); b"" at the end"""
        buf = np.empty(n, dtype=np.uint8)
        w = 0
        while n - w >= (1 << 16):
            got = self.lib.fc_bam_read_text(self.h, buf.ctypes.data + w, n - w)
            if got < 0:
                raise IOError("truncated or malformed BAM file (%d)" % got)
            if got == 0:
                break
            w += int(got)
        return buf[:w].tobytes()

    def _worker(self, n: int):
        try:
            while not self._stop:
                chunk = self._fill(n)
                self._queue.put(chunk)
                if not chunk:
                    return
        except Exception as e:  # handed to the reader
            self._queue.put(e)

    def read(self, n: int) -> bytes:
        if self._done:
            return b""
        n = max(int(n), 1 << 17)
        if self._prefetch <= 0:
            chunk = self._fill(n)
        else:
            if self._thread is None:
                import queue
                import threading

                self._queue = queue.Queue(maxsize=self._prefetch)
                self._thread = threading.Thread(target=self._worker, args=(n,), daemon=True)
                self._thread.start()
            chunk = self._queue.get()
            if isinstance(chunk, Exception):
                self._done = True
                raise chunk
        if not chunk:
            self._done = True
        return chunk

    def close(self):
        self._stop = True
        if self._thread is not None:
            while self._thread.is_alive():  # unblock a worker that waits for room in the queue
                try:
                    self._queue.get_nowait()
                except Exception:
                    pass
                self._thread.join(timeout=0.05)
            self._thread = None
        if self.h:
            self.lib.fc_bam_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
