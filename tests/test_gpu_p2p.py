"""
The multi-GPU kernels on ONE GPU (pytest -m gpu): scan_emit_p2p_kernel / emit_p2p_kernel / p2p_barrier_kernel and the
sliced accumulate path, driven through the C ABI with the ranks as contexts of this process on cuda:0
(fc_p2p_export_local / fc_p2p_connect_local: plain device pointers instead of CUDA IPC handles, a barrier that publishes
the slice counts without waiting -- kernels of different ranks must never spin on one another on a single device).

What a step has to deliver (find_circ.py:681-690, 584-590, 684-686 across ranks): the union of the owners' junction tables
is byte for byte the table one context builds from the union of all shards.
"""
import os

import numpy as np
import pytest

import helpers as H
from find_circ2_b200 import synth

pytestmark = pytest.mark.gpu


def _shard(g, J, rank, n, seed):
    t = synth.make_pairs(g, J, n, read_len=100, asize=20, seed=seed + rank, error_rate=0.01)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    return t, chrom, a_start, b_end, l, flags, internal


class Rank(object):
    """one context + its shard, device resident"""

    def __init__(self, torch, g, J, rank, n, seed, den=None):
        from find_circ2_b200.engine import Engine

        self.torch = torch
        dev = torch.device("cuda:0")
        tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        self.e = e = Engine(device=0, asize=20)
        e.load_genome_arrays(g.names, g.seqs)
        self.n = n
        if n == 0:
            z = np.zeros(0, np.int32)
            self.pairs = e.make_pairs(0, tn(z), tn(z), tn(z), tn(z), tn(np.zeros(0, np.uint8)), torch.zeros(3, dtype=torch.int32, device=dev), 1, 0)
            self.hits = torch.zeros(4, dtype=torch.int32, device=dev)
            z8 = tn(np.zeros(1, np.int64))
            self.pay = (tn(np.zeros(1, np.uint8)), tn(np.zeros(1, np.int16)), tn(np.zeros(1, np.int16)), z8, z8)
            self.cols = None
            return
        t, chrom, a_start, b_end, l, flags, internal = _shard(g, J, rank, n, seed)
        rh = e.hash_reads(t.reads, np.full(n, 100, dtype=np.int32))
        # read names repeat inside a shard (mates) and never across shards
        qh = ((t.name_id // 2).astype(np.uint64) + np.uint64(rank << 32)) * np.uint64(0x9E3779B97F4A7C15)
        qa = (t.as_a - np.maximum(t.xs_a, 0)).astype(np.int16)
        qb = (t.as_b - np.maximum(t.xs_b, 0)).astype(np.int16)
        wden = np.ones(n, np.uint8) if den is None else den(n)
        self.keep = []
        d = [tn(x) for x in (chrom, a_start, b_end, l, flags, internal)]
        n_words = Engine.n_words_for(int(l.max()))
        planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
        e.pack_reads(d[5], internal.shape[1], d[3], n_words, planes, d[4], 0)
        self.pairs = e.make_pairs(n, d[0], d[1], d[2], d[3], d[4], planes, n_words, int(l.max()))
        self.hits = torch.zeros(n * 4, dtype=torch.int32, device=dev)
        self.pay = (tn(wden), tn(qa), tn(qb), tn(rh.view(np.int64)), tn(qh.view(np.int64)))
        self.cols = (d[0], d[4])
        self.keep = [d, planes]


def _connect(ranks, capacity):
    world = len(ranks)
    ptrs = [r.e.p2p_export_local(capacity) for r in ranks]
    for k, r in enumerate(ranks):
        r.e.p2p_connect_local(world, k, [p[0] for p in ptrs], [p[1] for p in ptrs], np.full(world, capacity, np.int64))


def _reference_table(torch, g, ranks, bases):
    from find_circ2_b200.engine import Engine

    e = Engine(device=0, asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    e.agg_reset()
    for r, base in zip(ranks, bases):
        if r.n == 0:
            continue
        # same device arrays, another context: pointers are valid process-wide
        e.scan_emit(r.pairs, torch.zeros_like(r.hits), *r.pay, base, 0)
    t = e.agg_fetch(e.agg_finalize(0))
    e.close()
    return t


def _step(ranks, bases, total, fused=True, declare=True):
    for r in ranks:
        r.e.agg_reset_async(0)
        if declare:
            r.e.agg_set_idx_range(0, total)
    for r, base in zip(ranks, bases):
        if fused or r.n == 0:
            r.e.scan_emit_p2p(r.pairs, r.hits, *r.pay, base, 0)
        else:
            r.e.scan(r.pairs, r.hits, 0)
            r.e.agg_emit_p2p(r.n, r.hits, r.cols[0], r.cols[1], *r.pay, base, 0)
    for r in ranks:
        r.e.p2p_barrier(0)
    tables = [r.e.agg_fetch(r.e.agg_finalize(0)) for r in ranks]
    allj = np.concatenate(tables)
    return tables, allj[np.argsort(allj["first_idx"], kind="stable")]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def case():
    g = synth.make_genome([400000, 300000], seed=1, n_frac=0.005, n_run=(20, 300))
    J = synth.plant_junctions(g, 300, 60, seed=2, span=(150, 20000), margin=300)
    return g, J


@pytest.mark.parametrize("world,sizes", [(2, (30000, 27000)), (3, (20000, 0, 9000)), (8, (4000,) * 8)])
def test_peer_emit_equals_single_context(torch_cuda, case, world, sizes):
    """owners' tables == one context over the union; several steps in a row (the two buffer halves alternate); an empty
    shard still reduces the keys it owns"""
    g, J = case
    ranks = [Rank(torch_cuda, g, J, k, sizes[k], seed=40) for k in range(world)]
    bases = [int(sum(sizes[:k])) for k in range(world)]
    total = int(sum(sizes))
    _connect(ranks, int(2.5 * max(sizes)) + 4096)
    want = _reference_table(torch_cuda, g, ranks, bases)
    assert len(want) > 100
    for it, (fused, declare) in enumerate([(True, True), (False, True), (True, False), (True, True)]):
        tables, got = _step(ranks, bases, total, fused=fused, declare=declare)
        assert got.tobytes() == want.tobytes(), "step %d (fused=%s, declared range=%s)" % (it, fused, declare)
        # every junction key lives on exactly one rank
        keys = [set(zip(t["chrom"].tolist(), t["start"].tolist(), t["end"].tolist(), (t["sk"] & 3).tolist())) for t in tables]
        assert sum(len(k) for k in keys) == len(set().union(*keys)) == len(want)
        if world > 1:
            assert sum(1 for t in tables if len(t)) > 1
        assert sum(r.e.agg_n_records() for r in ranks) == int(want["n_spanned"].sum())
    for r in ranks:
        r.e.close()


def test_peer_emit_sort_based_path(torch_cuda, case, monkeypatch):
    """weights 1/3 and 1/5 (reads with 4 or 6 segments) send the step through the sort-based reduce, which first gathers the
    slices into one run; FC_AGG_MODE=sort forces that path for power-of-two weights too"""
    g, J = case
    sizes = (12000, 11000)
    den = lambda n: np.where(np.arange(n) % 7 == 0, 3, np.where(np.arange(n) % 11 == 0, 5, 1)).astype(np.uint8)  # noqa: E731
    ranks = [Rank(torch_cuda, g, J, k, sizes[k], seed=60, den=den) for k in range(2)]
    bases = [0, sizes[0]]
    _connect(ranks, 40000)
    want = _reference_table(torch_cuda, g, ranks, bases)
    assert (want["n_weighted"] != want["n_spanned"]).any()
    _, got = _step(ranks, bases, sum(sizes))
    assert got.tobytes() == want.tobytes()
    monkeypatch.setenv("FC_AGG_MODE", "sort")
    _, got = _step(ranks, bases, sum(sizes))
    assert got.tobytes() == want.tobytes()
    for r in ranks:
        r.e.close()


def test_peer_emit_protocol_errors(torch_cuda, case):
    from find_circ2_b200._lib import FindCircError

    g, J = case
    ranks = [Rank(torch_cuda, g, J, k, 6000, seed=80) for k in range(2)]
    # a slice of 500 records cannot hold what a rank sends to one owner: both sides of the step must fail loudly
    _connect(ranks, 1000)
    for r in ranks:
        r.e.agg_reset_async(0)
    for k, r in enumerate(ranks):
        r.e.scan_emit_p2p(r.pairs, r.hits, *r.pay, 6000 * k, 0)
    with pytest.raises(FindCircError) as ei:
        ranks[0].e.agg_finalize(0)  # before the barrier that ends the step
    assert ei.value.code == -8
    for r in ranks:
        r.e.p2p_barrier(0)
    for r in ranks:
        with pytest.raises(FindCircError) as ei:
            r.e.agg_finalize(0)
        assert ei.value.code == -7
    for r in ranks:
        r.e.close()
