"""
ctypes binding of libfindcirc_b200.so (include/findcirc_b200.h).

There is no fallback: if the library has not been built, or no CUDA device is present when a context is
created, an exception is raised.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C find_circ2_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfindcirc_b200.so")

FC_PF_BACKSPLICE = 1
FC_PF_MINUS = 2
FC_PF_READ_N = 4
FC_GENOME_PAD = 4096

ERRORS = {
    -1: "FC_E_CUDA", -2: "FC_E_ARG", -3: "FC_E_IO", -4: "FC_E_FORMAT", -5: "FC_E_NOGENOME", -6: "FC_E_RANGE",
    -7: "FC_E_NOMEM", -8: "FC_E_STATE", -9: "FC_E_COLLISION",
}


class FindCircError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "FC_E_?"), code, msg))
        self.code = code


class ScanParams(C.Structure):
    _fields_ = [
        ("asize", C.c_int32), ("margin", C.c_int32), ("maxdist", C.c_int32), ("noncanonical", C.c_int32),
        ("strandpref", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


class Pairs(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("d_chrom", C.c_void_p), ("d_a_start", C.c_void_p), ("d_b_end", C.c_void_p),
        ("d_l", C.c_void_p), ("d_flags", C.c_void_p), ("d_rlo", C.c_void_p), ("d_rhi", C.c_void_p), ("d_rn", C.c_void_p),
        ("n_words", C.c_int32), ("max_l", C.c_int32), ("plane_stride", C.c_int64),
    ]


class Batch(C.Structure):
    """fc_batch: the packed layout the scan kernels read"""
    _fields_ = [("n", C.c_int64), ("d_meta", C.c_void_p), ("d_reads", C.c_void_p), ("d_rn", C.c_void_p), ("n_words", C.c_int32),
                ("max_l", C.c_int32)]


class HostBatch(C.Structure):
    """fc_host_batch: one batch in fc_batch layout in (pinned) host memory, for fc_stream_submit"""
    _fields_ = [("n", C.c_int64), ("meta", C.c_void_p), ("reads", C.c_void_p), ("rn_idx", C.c_void_p), ("rn_rows", C.c_void_p),
                ("n_rn", C.c_int64), ("q", C.c_void_p), ("read_hash", C.c_void_p), ("qname_hash", C.c_void_p), ("idx", C.c_void_p),
                ("idx_base", C.c_uint64), ("n_words", C.c_int32), ("max_l", C.c_int32), ("emit", C.c_int32), ("out_mode", C.c_int32),
                ("out_hits", C.c_void_p), ("out_hit_mask", C.c_void_p), ("out_strand_mask", C.c_void_p)]


# numpy views of the C structs
HIT_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("w2", "<u4"), ("w3", "<u4")])
JREC_DTYPE = np.dtype(
    [
        ("chrom", "<u4"), ("start", "<i4"), ("end", "<i4"), ("sk", "<u4"), ("idx", "<u8"), ("read_hash", "<u8"),
        ("qname_hash", "<u8"), ("q_left", "<i2"), ("q_right", "<i2"), ("n_hits", "<u2"), ("dist", "u1"), ("ov", "u1"),
    ]
)
JUNCTION_DTYPE = np.dtype(
    [
        ("chrom", "<u4"), ("start", "<i4"), ("end", "<i4"), ("sk", "<u4"), ("first_idx", "<u8"), ("n_weighted", "<f8"),
        ("n_uniq_bridges", "<f8"), ("n_spanned", "<u4"), ("n_frags", "<u4"), ("n_uniq", "<u4"), ("best_q_left", "<i2"),
        ("best_q_right", "<i2"), ("min_n_hits", "<u2"), ("min_dist", "u1"), ("min_ov", "u1"), ("pad", "<u4"),
    ]
)
assert HIT_DTYPE.itemsize == 16 and JREC_DTYPE.itemsize == 48 and JUNCTION_DTYPE.itemsize == 64

# every symbol include/findcirc_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("fc_abi_version", C.c_int, []),
    ("fc_ctx_create", C.c_int, [C.c_int, C.POINTER(_P)]),
    ("fc_ctx_destroy", None, [_P]),
    ("fc_last_error", C.c_char_p, [_P]),
    ("fc_genome_load_fasta", C.c_int, [_P, C.c_char_p]),
    ("fc_genome_load_ascii", C.c_int, [_P, C.c_int32, _P, _P, _P]),
    ("fc_genome_share", C.c_int, [_P, _P]),
    ("fc_genome_n_chrom", C.c_int, [_P]),
    ("fc_genome_chrom_name", C.c_int, [_P, C.c_int32, C.c_char_p, C.c_int32]),
    ("fc_genome_chrom_size", C.c_int64, [_P, C.c_int32]),
    ("fc_genome_chrom_offset", C.c_int64, [_P, C.c_int32]),
    ("fc_genome_chrom_id", C.c_int, [_P, C.c_char_p]),
    ("fc_genome_stats", C.c_int, [_P, _P]),
    ("fc_genome_fetch", C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P]),
    ("fc_pack_reads", C.c_int, [_P, C.c_int64, _P, C.c_int32, _P, C.c_int32, _P, _P, _P, _P, _P]),
    ("fc_scan", C.c_int, [_P, C.POINTER(ScanParams), C.POINTER(Pairs), _P, _P]),
    ("fc_scan_emit", C.c_int, [_P, C.POINTER(ScanParams), C.POINTER(Pairs), _P, _P, _P, _P, _P, _P, C.c_uint64, _P, _P]),
    ("fc_batch_words", C.c_int32, [C.c_int32]),
    ("fc_batch_pack", C.c_int, [_P, C.POINTER(Pairs), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("fc_scan_batch", C.c_int, [_P, C.POINTER(ScanParams), _P, _P, _P]),
    ("fc_scan_emit_batch", C.c_int, [_P, C.POINTER(ScanParams), _P, _P, _P, _P, _P, C.c_uint64, _P, _P]),
    ("fc_stream_create", C.c_int, [_P, C.c_int32, C.c_int64, C.c_int32, C.POINTER(_P)]),
    ("fc_stream_destroy", None, [_P]),
    ("fc_stream_submit", C.c_int, [_P, C.c_int32, C.POINTER(ScanParams), _P]),
    ("fc_stream_wait", C.c_int, [_P, C.c_int32]),
    ("fc_stream_query", C.c_int, [_P, C.c_int32]),
    ("fc_scan_ties", C.c_int, [_P, C.POINTER(ScanParams), C.POINTER(Pairs), _P, _P, _P, _P]),
    ("fc_scan_host", C.c_int, [_P, C.POINTER(ScanParams), C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    ("fc_batch_host", C.c_int, [_P, C.POINTER(ScanParams), C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P,
                                C.c_uint64, C.c_int32, _P]),
    ("fc_batch_host_idx", C.c_int, [_P, C.POINTER(ScanParams), C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P,
                                    C.c_uint64, C.c_int32, _P]),
    ("fc_batch_host_planes", C.c_int, [_P, C.POINTER(ScanParams), C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int64,
                                       C.c_int32, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_int32, _P]),
    ("fc_agg_emit_idx", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("fc_ingest_create", _P, [_P, C.c_int32, _P, _P]),
    ("fc_ingest_destroy", None, [_P]),
    ("fc_ingest_set_position", C.c_int, [_P, C.c_int64, C.c_int32]),
    ("fc_ingest_position", C.c_int64, [_P]),
    ("fc_ingest_parse", C.c_int64, [_P, _P, C.c_int64, C.c_int32, _P]),
    ("fc_agg_reset", C.c_int, [_P]),
    ("fc_agg_emit", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_uint64, _P]),
    ("fc_batch_emit_host", C.c_int, [_P, _P, C.c_uint64]),
    ("fc_batch_ties_host", C.c_int, [_P, C.POINTER(ScanParams), _P, _P]),
    ("fc_agg_append", C.c_int, [_P, C.c_int64, _P, _P]),
    ("fc_agg_append_host", C.c_int, [_P, C.c_int64, _P]),
    ("fc_agg_replace", C.c_int, [_P, C.c_int64, _P, _P]),
    ("fc_agg_n_records", C.c_int64, [_P]),
    ("fc_agg_records", _P, [_P]),
    ("fc_agg_partition", C.c_int, [_P, C.c_int32, _P, _P, _P]),
    ("fc_agg_finalize", C.c_int64, [_P, _P]),
    ("fc_agg_fetch", C.c_int, [_P, C.c_int64, _P]),
    ("fc_agg_junctions", _P, [_P]),
    ("fc_p2p_export", C.c_int, [_P, C.c_int64, _P]),
    ("fc_p2p_connect", C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    ("fc_agg_emit_p2p", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_uint64, _P]),
    ("fc_scan_emit_p2p", C.c_int, [_P, C.POINTER(ScanParams), C.POINTER(Pairs), _P, _P, _P, _P, _P, _P, C.c_uint64, _P]),
    ("fc_p2p_export_local", C.c_int, [_P, C.c_int64, _P, _P]),
    ("fc_p2p_connect_local", C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P]),
    ("fc_p2p_set_timeout", C.c_int, [_P, C.c_double]),
    ("fc_p2p_barrier", C.c_int, [_P, _P]),
    ("fc_agg_reset_async", C.c_int, [_P, _P]),
    ("fc_bam_open", _P, [C.c_char_p]),
    ("fc_bam_close", None, [_P]),
    ("fc_bam_n_ref", C.c_int32, [_P]),
    ("fc_bam_ref_name", C.c_char_p, [_P, C.c_int32]),
    ("fc_bam_ref_length", C.c_int64, [_P, C.c_int32]),
    ("fc_bam_read_text", C.c_int64, [_P, _P, C.c_int64]),
    ("fc_ingest_evidence", C.c_int, [_P, _P]),
    ("fc_unique_rows", C.c_int64, [_P, C.c_int64, C.c_int32, _P, _P]),
    ("fc_text_gather", C.c_int64, [_P, C.c_int64, _P, _P, _P]),
    ("fc_fastq_format", C.c_int64, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int64, _P]),
    ("fc_merge_tables", C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    ("fc_pinned_alloc", _P, [C.c_int64]),
    ("fc_pinned_free", None, [_P]),
    ("fc_device_sync", C.c_int, [_P]),
    ("fc_agg_set_idx_range", C.c_int, [_P, C.c_uint64, C.c_uint64]),
    ("fc_agg_set_timing", C.c_int, [_P, C.c_int32]),
    ("fc_agg_get_timing", C.c_int, [_P, _P]),
    ("fc_launch_count", C.c_int64, [_P]),
    ("fc_hash_bytes", C.c_uint64, [_P, C.c_int64]),
    ("fc_hash_read", C.c_uint64, [_P, C.c_int64, _P]),
    ("fc_hash_reads_host", C.c_int, [C.c_int64, _P, C.c_int32, _P, _P, _P]),
    ("fc_hash_reads_device", C.c_int, [_P, C.c_int64, _P, C.c_int32, _P, C.c_int32, _P, _P]),
]

_lib = None


def load():
    """dlopen the library and declare every prototype (raises if the library is missing)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build the CUDA library first (make -C find_circ2_b200/csrc); "
                "there is no CPU fallback" % LIB_PATH
            )
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def ptr(a):
    """device/host pointer of a numpy array, torch tensor, int or None"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor
