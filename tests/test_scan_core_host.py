"""
The DEVICE scan code (find_circ2_b200/csrc/scan_core.cuh) compiled for the host must agree with the oracle's
per-split loop (find_circ.py:904-974) on random pairs: bit-parallel path, per-base path, every option.
This is a development check that runs without a GPU; the GPU parity tests (-m gpu) repeat it through the C ABI.
"""
import numpy as np
import pytest

from find_circ2_b200 import synth
from oracle import find_circ_oracle as O

import helpers as H


def _case(read_len, asize, seed, n=1500, error_rate=0.01):
    g = synth.make_genome([40000, 30000, 5000], seed=seed, n_frac=0.01, n_run=(20, 200))
    J = synth.plant_junctions(g, 60, 40, seed=seed + 1, span=(150, 8000), margin=300)
    t = synth.make_pairs(g, J, n, read_len=read_len, asize=asize, seed=seed + 2, error_rate=error_rate, frac_decoy=0.15,
                         frac_edge=0.05, frac_read_n=0.03, frac_inner_shift=0.2)
    return g, t


@pytest.mark.parametrize(
    "read_len,asize,margin,maxdist,nonc,spref",
    [
        (100, 20, 2, 2, 0, 0),
        (100, 15, 2, 2, 0, 0),
        (76, 20, 2, 2, 0, 0),
        (150, 20, 2, 2, 0, 0),
        (250, 20, 2, 2, 0, 0),
        (100, 20, 0, 0, 0, 0),
        (100, 20, 4, 3, 0, 1),
        (100, 20, 2, 2, 1, 0),
        (100, 20, 2, 2, 1, 1),
        (44, 20, 2, 2, 0, 0),
        (36, 20, 2, 2, 0, 0),
    ],
)
def test_host_compiled_scan_matches_oracle(read_len, asize, margin, maxdist, nonc, spref):
    g, t = _case(read_len, asize, seed=100 + read_len + asize + margin)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, asize, margin)
    opt = O.Options(asize=asize, margin=margin, maxdist=maxdist, noncanonical=bool(nonc), strandpref=bool(spref))
    want = H.oracle_scan(H.GenomeStrings(g), g.names, chrom, a_start, b_end, l, flags, internal, opt)
    for mode in (0, 1, 2):
        got = H.harness_scan(g, chrom, a_start, b_end, l, flags, internal, margin, maxdist, nonc, spref, mode=mode)
        dec = [H.decode_hit(r) for r in got]
        bad = [(i, dec[i], want[i]) for i in range(len(want)) if dec[i] != want[i]]
        assert not bad, (mode, bad[:5])
    n_hit = sum(1 for w in want if w)
    assert n_hit > 0.3 * len(want) or read_len < 50


def test_tile_store_built_for_longer_windows():
    """a tile store built for longer reads (other T / stride) must give the same answers on short reads"""
    g, t = _case(100, 20, seed=7)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    base = H.harness_scan(g, chrom, a_start, b_end, l, flags, internal, 2, 2)
    for w in (80, 96, 118, 128, 200, 256):
        other = H.harness_scan(g, chrom, a_start, b_end, l, flags, internal, 2, 2, tile_window=w)
        assert np.array_equal(base[:, :3], other[:, :3]), w
    # tile store too small for the batch -> master planes, same answers
    narrow = H.harness_scan(g, chrom, a_start, b_end, l, flags, internal, 2, 2, tile_window=40)
    assert np.array_equal(base[:, :3], narrow[:, :3])
