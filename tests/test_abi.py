"""the C-ABI library loads and exports every symbol include/findcirc_b200.h declares (no GPU needed)"""
import os
import re

import numpy as np

from conftest import ROOT
from find_circ2_b200 import _lib


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "findcirc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.fc_abi_version() == 2


def test_struct_sizes_match_header():
    assert _lib.HIT_DTYPE.itemsize == 16
    assert _lib.JREC_DTYPE.itemsize == 48
    assert _lib.JUNCTION_DTYPE.itemsize == 64


def test_context_creation_fails_loudly_without_gpu():
    import ctypes as C

    import torch

    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.fc_ctx_create(0, C.byref(h))
    assert rc < 0
    assert b"no CPU path" in lib.fc_last_error(None)


def test_read_hash_is_strand_invariant():
    lib = _lib.load()
    a = b"ACGTTGCAAGGCTTAN"
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    rc = a.translate(comp)[::-1]
    ha, hb = lib.fc_hash_read(a, len(a), None), lib.fc_hash_read(rc, len(rc), None)
    assert ha == hb and (ha & 1) == 0
    pal = b"ACGT"  # its own reverse complement
    assert lib.fc_hash_read(pal, 4, None) & 1 == 1
    assert lib.fc_hash_read(b"ACGA", 4, None) != lib.fc_hash_read(b"ACGC", 4, None)
    seqs = np.frombuffer(a + rc, dtype=np.uint8).reshape(2, -1)
    out = np.zeros(2, dtype=np.uint64)
    lens = np.array([len(a), len(a)], dtype=np.int32)
    assert lib.fc_hash_reads_host(2, seqs.ctypes.data, seqs.shape[1], lens.ctypes.data, out.ctypes.data, None) == 0
    assert out[0] == out[1] == ha
