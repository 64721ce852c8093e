"""shared helpers of the test-suite (test infrastructure; may use the oracle)"""
import ctypes
import os
import subprocess

import numpy as np

from oracle import find_circ_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIG_LETTERS = "ACGTN"


def decode_sig(code):
    return "".join(SIG_LETTERS[(code >> (3 * k)) & 7] for k in range(4))


def decode_hit(h):
    """(start, end, strand, dist, ov, signal, n_hits) or None from one fc_hit row (4 x uint32 view)"""
    start, end, w2, w3 = int(h[0]), int(h[1]), int(h[2]), int(h[3])
    n_hits = w2 & 0xFFFF
    if n_hits == 0:
        return None
    start = start - (1 << 32) if start >= (1 << 31) else start
    end = end - (1 << 32) if end >= (1 << 31) else end
    return (start, end, "-" if (w3 & 1) else "+", (w2 >> 16) & 0xFF, w2 >> 24, decode_sig((w3 >> 1) & 0xFFF), n_hits)


def pairs_to_soa(t, asize, margin):
    """PairTable (find_circ2_b200.synth) -> scan inputs for 2-segment reads that tile the read (q_start=0, q_end=R)"""
    eff = asize - margin
    R = t.read_len
    n = len(t)
    a_start = (t.a_pos + eff).astype(np.int32)
    b_end = (t.b_pos + t.b_len - eff).astype(np.int32)
    l = np.full(n, R - 2 * eff, dtype=np.int32)
    backsplice = (t.b_pos - (t.a_pos + t.a_len)) < 0
    flags = (backsplice.astype(np.uint8) * 1) | (t.reverse.astype(np.uint8) * 2)
    internal = np.ascontiguousarray(t.reads[:, eff : R - eff])
    return t.chrom.astype(np.int32), a_start, b_end, l, flags.astype(np.uint8), internal


class GenomeStrings:
    """oracle-style genome.get over a synth genome"""

    def __init__(self, g):
        self.seqs = {n: s.tobytes().decode() for n, s in zip(g.names, g.seqs)}
        self.names = list(g.names)

    get = O.Genome.get


def oracle_scan(gs, names, chrom, a_start, b_end, l, flags, internal, opt):
    """first tie + n_hits per pair, through the oracle's per-split loop"""
    out = []
    for i in range(len(chrom)):
        li = int(l[i])
        if li < 0:
            out.append(None)
            continue
        c = names[int(chrom[i])]
        a0, b1 = int(a_start[i]), int(b_end[i])
        af = gs.get(c, a0, a0 + li + 2).upper()
        bf = gs.get(c, b1 - li - 2, b1).upper()
        inner = internal[i, :li].tobytes().decode().upper()
        hits = O.scan_windows(af, bf, inner, c, a0, b1, bool(flags[i] & 1), "-" if flags[i] & 2 else "+", opt)
        if not hits:
            out.append(None)
        else:
            h = hits[0]
            out.append((h.start, h.end, h.strand, int(h.dist), h.ov, h.gtag, h.n_hits))
    return out


# ---------------------------------------------------------------------------------------------- host harness
_HARNESS = None


def host_harness():
    """device scan code compiled for the host (tests/tools/scan_host_harness.cpp)"""
    global _HARNESS
    if _HARNESS is None:
        src = os.path.join(ROOT, "tests", "tools", "scan_host_harness.cpp")
        so = os.path.join(ROOT, "tests", "tools", "libscan_host_harness.so")
        core = os.path.join(ROOT, "find_circ2_b200", "csrc", "scan_core.cuh")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
            subprocess.check_call(
                ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, src]
            )
        lib = ctypes.CDLL(so)
        lib.hh_genome_build.restype = ctypes.c_void_p
        lib.hh_genome_build.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.hh_genome_free.argtypes = [ctypes.c_void_p]
        lib.hh_build_tiles.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.hh_pack_reads.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [
            ctypes.c_void_p
        ] * 4
        lib.hh_scan.argtypes = (
            [ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_int64] + [ctypes.c_void_p] * 8 + [ctypes.c_int, ctypes.c_void_p]
        )
        _HARNESS = lib
    return _HARNESS


def harness_scan(g, chrom, a_start, b_end, l, flags, internal, margin, maxdist, noncanonical=0, strandpref=0, mode=0,
                 tile_window=None):
    """mode 0: tile store + kernel choice as fc_scan; 1: master planes only; 2: per-base path"""
    lib = host_harness()
    seqs = [np.ascontiguousarray(s) for s in g.seqs]
    ptrs = (ctypes.c_void_p * len(seqs))(*[s.ctypes.data for s in seqs])
    sizes = np.array([len(s) for s in seqs], dtype=np.int64)
    h = lib.hh_genome_build(len(seqs), ptrs, sizes.ctypes.data)
    try:
        n = len(chrom)
        max_l = int(max(l.max(), 0)) if n else 0
        lib.hh_build_tiles(h, tile_window if tile_window is not None else max_l + 2)
        n_words = max(1, (max_l + 31) // 32)
        stride = internal.shape[1] if internal.ndim == 2 and internal.shape[1] else 1
        internal = np.ascontiguousarray(internal)
        planes = np.zeros((3, n_words * n), dtype=np.uint32)
        flags = flags.copy()
        lib.hh_pack_reads(n, internal.ctypes.data, stride, l.ctypes.data, n_words, planes[0].ctypes.data, planes[1].ctypes.data,
                          planes[2].ctypes.data, flags.ctypes.data)
        out = np.zeros((n, 4), dtype=np.uint32)
        rc = lib.hh_scan(h, margin, maxdist, noncanonical, strandpref, mode, max_l, n, chrom.ctypes.data, a_start.ctypes.data,
                         b_end.ctypes.data, l.ctypes.data, flags.ctypes.data, planes[0].ctypes.data, planes[1].ctypes.data,
                         planes[2].ctypes.data, n_words, out.ctypes.data)
        assert rc == 0
        return out
    finally:
        lib.hh_genome_free(h)


def compare_outputs(circ, lin, reads, multi, counters, ref_dir, argv, test_results=None):
    """the five outputs of a run against a reference run stored under tests/golden/<case>/ref_*; an output that the run
    redirects with --stdout is compared with the captured stdout and its file holds one comment line (find_circ.py:453-458)"""
    rd = lambda n: open(os.path.join(ref_dir, n)).read()  # noqa: E731
    redirected = argv[argv.index("--stdout") + 1] if "--stdout" in argv else ""
    files = {"circs": "circ_splice_sites.bed", "lins": "lin_splice_sites.bed", "reads": "spliced_reads.fastq", "multi": "multi_events.tsv"}
    got = {"circs": circ, "lins": lin, "reads": reads, "multi": multi}
    canon = {"circs": O.canonical_bed, "lins": O.canonical_bed, "reads": lambda t: t, "multi": O.canonical_multi}
    for name, fn in files.items():
        if name == redirected:
            want = rd("stdout.txt")
            if name != "reads":  # (the reads file is a gzip stream the reference leaves truncated in this case)
                assert rd(fn) == "# redirected to stdout\n"
        else:
            want = rd(fn)
        assert canon[name](got[name]) == canon[name](want), name
    assert counters == rd("counters.txt")
    if "--test" in argv:  # test_results.tsv: one row per fragment that reached record_hits, in stream order
        assert test_results == rd("test_results.tsv")
