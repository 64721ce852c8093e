"""
GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libfindcirc_b200.so; the oracle (oracle/find_circ_oracle.py) is the checker.  Bit-exact.
"""
import os
from collections import defaultdict

import numpy as np
import pytest

import helpers as H
from conftest import GOLDEN
from find_circ2_b200 import synth
from oracle import find_circ_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


def _engine(**kw):
    from find_circ2_b200.engine import Engine

    return Engine(device=0, **kw)


def _case(read_len, asize, seed, n=3000, error_rate=0.01, sizes=(60000, 45000, 7000)):
    g = synth.make_genome(list(sizes), seed=seed, n_frac=0.01, n_run=(20, 200))
    J = synth.plant_junctions(g, 80, 50, seed=seed + 1, span=(150, 8000), margin=300)
    t = synth.make_pairs(g, J, n, read_len=read_len, asize=asize, seed=seed + 2, error_rate=error_rate, frac_decoy=0.15,
                         frac_edge=0.05, frac_read_n=0.03, frac_inner_shift=0.2)
    return g, J, t


# ---------------------------------------------------------------------------------------------- genome store
def test_genome_fetch_matches_reference_get(tmp_path):
    """device 2-bit + N-mask store vs indexed_fasta.get_data semantics (find_circ.py:189-215), incl. N padding,
    soft-masked letters and ragged last FASTA lines"""
    fa = os.path.join(GOLDEN, "synth_a", "genome.fa")
    og = O.Genome(fa)
    e = _engine()
    e.load_genome_fasta(fa)
    assert e.chrom_names == og.names
    assert e.chrom_sizes == [og.size(n) for n in og.names]
    rng = np.random.default_rng(5)
    for name in og.names:
        cid = e.chrom_id(name)
        size = og.size(name)
        assert e.fetch(cid, 0, size) == og.get(name, 0, size).upper()
        for _ in range(50):
            s = int(rng.integers(-300, size + 100))
            ln = int(rng.integers(1, 400))
            assert e.fetch(cid, s, s + ln) == og.get(name, s, s + ln).upper(), (name, s, ln)
        assert e.fetch(cid, -100, 5) == "N" * 100 + og.get(name, 0, 5).upper()
        assert e.fetch(cid, size - 3, size + 50) == og.get(name, size - 3, size).upper() + "N" * 50
    st = e.genome_stats()
    assert st["bases"] == sum(e.chrom_sizes)
    assert st["n"] == sum(og.seqs[n].upper().count("N") for n in og.names)
    with pytest.raises(KeyError):
        e.chrom_id("no_such_chrom")
    e.close()


def test_genome_from_reference_fixtures():
    for case in ("kat3", "cdr1as"):
        fa = os.path.join(GOLDEN, case, "genome.fa")
        og = O.Genome(fa)
        e = _engine()
        e.load_genome_fasta(fa)
        for name in og.names:
            assert e.fetch(e.chrom_id(name), -10, og.size(name) + 10) == og.get(name, -10, og.size(name) + 10).upper()
        e.close()


def test_missing_genome_is_an_error(tmp_path):
    from find_circ2_b200._lib import FindCircError

    e = _engine()
    with pytest.raises(FindCircError):
        e.load_genome_fasta(str(tmp_path / "nope.fa"))
    with pytest.raises(FindCircError):
        e.scan_host(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32),
                    np.zeros(1, np.uint8), np.zeros((1, 4), np.uint8))
    e.close()


# ---------------------------------------------------------------------------------------------- scan
PARAM_SETS = [
    (100, 20, 2, 2, 0, 0),
    (100, 15, 2, 2, 0, 0),
    (76, 20, 2, 2, 0, 0),
    (150, 20, 2, 2, 0, 0),
    (250, 20, 2, 2, 0, 0),
    (300, 15, 2, 2, 0, 0),
    (100, 20, 0, 0, 0, 0),
    (100, 20, 4, 3, 0, 1),
    (100, 20, 2, 2, 1, 0),
    (100, 20, 2, 2, 1, 1),
    (44, 20, 2, 2, 0, 0),
    (36, 20, 2, 2, 0, 0),
]


@pytest.mark.parametrize("read_len,asize,margin,maxdist,nonc,spref", PARAM_SETS)
def test_scan_matches_oracle(read_len, asize, margin, maxdist, nonc, spref):
    g, J, t = _case(read_len, asize, seed=300 + read_len + asize + margin + nonc)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, asize, margin)
    opt = O.Options(asize=asize, margin=margin, maxdist=maxdist, noncanonical=bool(nonc), strandpref=bool(spref))
    want = H.oracle_scan(H.GenomeStrings(g), g.names, chrom, a_start, b_end, l, flags, internal, opt)
    e = _engine(asize=asize, margin=margin, maxdist=maxdist, noncanonical=bool(nonc), strandpref=bool(spref))
    e.load_genome_arrays(g.names, g.seqs)
    hits = e.scan_host(chrom, a_start, b_end, l, flags, internal)
    got = [H.decode_hit(r) for r in hits.view(np.uint32).reshape(-1, 4)]
    bad = [(i, got[i], want[i]) for i in range(len(want)) if got[i] != want[i]]
    assert not bad, bad[:5]
    assert sum(1 for w in want if w) > 0.3 * len(want) or read_len < 50
    e.close()


def test_scan_edge_inputs():
    """empty batch, l < 0, l == 0, ragged lengths inside one batch"""
    g, J, t = _case(100, 20, seed=77, n=400)
    e = _engine(asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    z = np.zeros(0, np.int32)
    assert len(e.scan_host(z, z, z, z, np.zeros(0, np.uint8), np.zeros((0, 8), np.uint8))) == 0
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    # ragged: shorten the internal part of some pairs (as shorter reads in the same batch would)
    l = l.copy()
    a_start = a_start.copy()
    rng = np.random.default_rng(3)
    cut = rng.integers(0, 70, size=len(l))
    l = np.maximum(l - cut, -3).astype(np.int32)
    opt = O.Options(asize=20)
    want = H.oracle_scan(H.GenomeStrings(g), g.names, chrom, a_start, b_end, l, flags, internal, opt)
    hits = e.scan_host(chrom, a_start, b_end, l, flags, internal)
    got = [H.decode_hit(r) for r in hits.view(np.uint32).reshape(-1, 4)]
    assert got == want
    assert (l == 0).any() or True
    e.close()


def test_device_path_equals_host_path(torch_cuda):
    torch = torch_cuda
    g, J, t = _case(100, 20, seed=9)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    e = _engine(asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    host = e.scan_host(chrom, a_start, b_end, l, flags, internal)
    dev = torch.device("cuda:0")
    n = len(chrom)
    n_words = max(1, (int(l.max()) + 31) // 32)
    d = {k: torch.from_numpy(v).to(dev) for k, v in dict(chrom=chrom, a=a_start, b=b_end, l=l, fl=flags, asc=internal).items()}
    planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    out = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    e.pack_reads(d["asc"], internal.shape[1], d["l"], n_words, planes, d["fl"], st)
    pairs = e.make_pairs(n, d["chrom"], d["a"], d["b"], d["l"], d["fl"], planes, n_words, int(l.max()))
    e.scan(pairs, out, st)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint32).reshape(-1, 4)
    assert np.array_equal(got, host.view(np.uint32).reshape(-1, 4))
    e.close()


def test_all_ties_enumeration(torch_cuda):
    """--all-hits: every tie in the reference's order (find_circ.py:966-974, 1312-1317)"""
    torch = torch_cuda
    g, J, t = _case(100, 20, seed=21, n=1200)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    for nonc in (1, 0):
        opt = O.Options(asize=20, noncanonical=bool(nonc), allhits=True)
        e = _engine(asize=20, noncanonical=bool(nonc))
        e.load_genome_arrays(g.names, g.seqs)
        dev = torch.device("cuda:0")
        n = len(chrom)
        n_words = max(1, (int(l.max()) + 31) // 32)
        d = {k: torch.from_numpy(v).to(dev) for k, v in dict(chrom=chrom, a=a_start, b=b_end, l=l, fl=flags, asc=internal).items()}
        planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
        out = torch.zeros(n * 4, dtype=torch.int32, device=dev)
        e.pack_reads(d["asc"], internal.shape[1], d["l"], n_words, planes, d["fl"], 0)
        pairs = e.make_pairs(n, d["chrom"], d["a"], d["b"], d["l"], d["fl"], planes, n_words, int(l.max()))
        e.scan(pairs, out, 0)
        torch.cuda.synchronize()
        hits = out.cpu().numpy().view(np.uint32).reshape(-1, 4)
        nh = (hits[:, 2] & 0xFFFF).astype(np.int64)
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum(nh)
        d_off = torch.from_numpy(off).to(dev)
        ties = torch.zeros(max(int(off[-1]), 1) * 4, dtype=torch.int32, device=dev)
        e.scan_ties(pairs, out, d_off, ties, 0)
        torch.cuda.synchronize()
        tt = ties.cpu().numpy().view(np.uint32).reshape(-1, 4)
        gs = H.GenomeStrings(g)
        for i in range(n):
            li = int(l[i])
            c = g.names[int(chrom[i])]
            a0, b1 = int(a_start[i]), int(b_end[i])
            want = O.scan_windows(gs.get(c, a0, a0 + li + 2).upper(), gs.get(c, b1 - li - 2, b1).upper(),
                                  internal[i, :li].tobytes().decode(), c, a0, b1, bool(flags[i] & 1),
                                  "-" if flags[i] & 2 else "+", opt)
            want = [(h.start, h.end, h.strand, int(h.dist), h.ov, h.gtag, h.n_hits) for h in want]
            got = [H.decode_hit(tt[k]) for k in range(off[i], off[i + 1])]
            assert got == want, (i, got[:3], want[:3])
        e.close()


# ---------------------------------------------------------------------------------------------- aggregation
def _py_aggregate(recs):
    """dict-based restatement of Hit.add / reductions (find_circ.py:526-600) over fc_jrec-like rows"""
    acc = {}
    for r in recs:
        key = (int(r["chrom"]), int(r["start"]), int(r["end"]), int(r["sk"]) & 3)
        a = acc.setdefault(key, dict(first=int(r["idx"]), w=0.0, b=0.0, n=0, ql=[], qr=[], d=[], o=[], nh=[], rh=set(),
                                     qh=set(), pal=set(), sig=0))
        wt = 1.0 / ((int(r["sk"]) >> 8) & 0xFF)
        a["w"] += wt
        if r["q_left"] != 0 and r["q_right"] != 0:
            a["b"] += wt
        a["n"] += 1
        a["ql"].append(int(r["q_left"]))
        a["qr"].append(int(r["q_right"]))
        a["d"].append(int(r["dist"]))
        a["o"].append(int(r["ov"]))
        a["nh"].append(int(r["n_hits"]))
        a["rh"].add(int(r["read_hash"]))
        if int(r["read_hash"]) & 1:
            a["pal"].add(int(r["read_hash"]))
        a["qh"].add(int(r["qname_hash"]))
        a["sig"] = (int(r["sk"]) >> 16) & 0xFFF
    out = []
    for key, a in sorted(acc.items(), key=lambda kv: kv[1]["first"]):
        out.append((key, a["first"], a["w"], a["b"], a["n"], len(a["qh"]), len(a["rh"]) - (len(a["pal"]) + 1) // 2,
                    max(a["ql"]), max(a["qr"]), min(a["nh"]), min(a["d"]), min(a["o"]), a["sig"]))
    return out


def _junction_rows(j):
    return [((int(r["chrom"]), int(r["start"]), int(r["end"]), int(r["sk"]) & 3), int(r["first_idx"]), float(r["n_weighted"]),
             float(r["n_uniq_bridges"]), int(r["n_spanned"]), int(r["n_frags"]), int(r["n_uniq"]), int(r["best_q_left"]),
             int(r["best_q_right"]), int(r["min_n_hits"]), int(r["min_dist"]), int(r["min_ov"]), (int(r["sk"]) >> 16) & 0xFFF)
            for r in j]


def _random_records(n, n_keys, seed, dens=(1, 1, 1, 2, 2, 3, 4, 5, 8)):
    from find_circ2_b200._lib import JREC_DTYPE

    rng = np.random.default_rng(seed)
    keys = np.zeros(n_keys, dtype=[("chrom", "u4"), ("start", "u4"), ("end", "u4"), ("s", "u4")])
    keys["chrom"] = rng.integers(0, 5, n_keys)
    keys["start"] = rng.integers(0, 1 << 20, n_keys)
    keys["end"] = keys["start"] + rng.integers(1, 1 << 16, n_keys)
    keys["s"] = rng.integers(0, 4, n_keys)
    w = 1.0 / np.arange(1, n_keys + 1)
    pick = rng.choice(n_keys, size=n, p=w / w.sum())
    r = np.zeros(n, dtype=JREC_DTYPE)
    r["chrom"] = keys["chrom"][pick]
    r["start"] = keys["start"][pick]
    r["end"] = keys["end"][pick]
    den = rng.choice(np.array(dens), size=n)
    r["sk"] = keys["s"][pick] | (den.astype(np.uint32) << 8) | (np.uint32(0x4D3) << 16)
    r["idx"] = np.arange(n, dtype=np.uint64) * 3 + 7
    # few distinct reads / names so that duplicates occur; some palindromes (bit 0)
    r["read_hash"] = rng.integers(0, max(n // 3, 2), n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    r["qname_hash"] = rng.integers(0, max(n // 2, 2), n).astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
    r["q_left"] = rng.integers(-5, 60, n)
    r["q_right"] = rng.integers(-5, 60, n)
    r["q_left"][rng.random(n) < 0.1] = 0
    r["n_hits"] = rng.integers(1, 5, n)
    r["dist"] = rng.integers(0, 3, n)
    r["ov"] = rng.integers(0, 3, n)
    return r


@pytest.mark.parametrize("n,n_keys,seed", [(1, 1, 1), (37, 5, 2), (5000, 300, 3), (200000, 5000, 4), (300000, 3, 5)])
def test_aggregation_matches_python(n, n_keys, seed):
    recs = _random_records(n, n_keys, seed)
    e = _engine()
    e.agg_reset()
    # appended in three pieces: stream order must survive
    cuts = [0, n // 3, n // 2, n]
    for a, b in zip(cuts, cuts[1:]):
        e.agg_append_host(recs[a:b])
    assert e.agg_n_records() == n
    nj = e.agg_finalize()
    got = _junction_rows(e.agg_fetch(nj))
    want = _py_aggregate(recs)
    assert len(got) == len(want)
    for gr, wr in zip(got, want):
        assert gr == wr, (gr, wr)
    # finalize is idempotent
    assert e.agg_finalize() == nj
    e.agg_reset()
    assert e.agg_finalize() == 0
    e.close()


def test_emit_from_scan_matches_python(torch_cuda):
    """scan -> fc_agg_emit -> finalize on the device == python aggregation of the oracle's first ties"""
    torch = torch_cuda
    g, J, t = _case(100, 20, seed=31, n=6000, error_rate=0.005)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    n = len(chrom)
    opt = O.Options(asize=20)
    want_hits = H.oracle_scan(H.GenomeStrings(g), g.names, chrom, a_start, b_end, l, flags, internal, opt)
    e = _engine(asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(8)
    wden = rng.choice(np.array([1, 1, 1, 2, 3], dtype=np.uint8), size=n)
    q_a = (t.as_a - np.maximum(t.xs_a, 0)).astype(np.int16)
    q_b = (t.as_b - np.maximum(t.xs_b, 0)).astype(np.int16)
    lens = np.full(n, t.read_len, dtype=np.int32)
    rh = e.hash_reads(t.reads, lens)
    qh = np.array([e.hash_bytes(("r%d" % (i // 2)).encode()) for i in range(n)], dtype=np.uint64)
    n_words = max(1, (int(l.max()) + 31) // 32)
    tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_chrom, d_a, d_b, d_l, d_fl, d_asc = tn(chrom), tn(a_start), tn(b_end), tn(l), tn(flags), tn(internal)
    planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    out = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    e.pack_reads(d_asc, internal.shape[1], d_l, n_words, planes, d_fl, 0)
    pairs = e.make_pairs(n, d_chrom, d_a, d_b, d_l, d_fl, planes, n_words, int(l.max()))
    e.agg_reset()
    half = n // 2
    e.scan(pairs, out, 0)
    d_w, d_qa, d_qb, d_rh, d_qh = tn(wden), tn(q_a), tn(q_b), tn(rh.view(np.int64)), tn(qh.view(np.int64))
    # emit in two batches with consecutive idx ranges
    for lo, hi in ((0, half), (half, n)):
        e.agg_emit(hi - lo, out[4 * lo:], d_chrom[lo:], d_fl[lo:], d_w[lo:], d_qa[lo:], d_qb[lo:], d_rh[lo:], d_qh[lo:], 1000 + lo, 0)
    nj = e.agg_finalize()
    got = _junction_rows(e.agg_fetch(nj))
    # python side
    from find_circ2_b200._lib import JREC_DTYPE

    rows = []
    for i, h in enumerate(want_hits):
        if h is None:
            continue
        start, end, strand, dist, ov, sig, nh = h
        back = bool(flags[i] & 1)
        r = np.zeros((), dtype=JREC_DTYPE)
        r["chrom"], r["start"], r["end"] = chrom[i], start, end
        sigc = sum("ACGTN".index(ch) << (3 * k) for k, ch in enumerate(sig))
        r["sk"] = (1 if strand == "-" else 0) | (0 if back else 2) | (int(rh[i]) & 1) << 2 | int(wden[i]) << 8 | sigc << 16
        r["idx"] = 1000 + i
        r["read_hash"], r["qname_hash"] = rh[i], qh[i]
        r["q_left"], r["q_right"] = (q_b[i], q_a[i]) if back else (q_a[i], q_b[i])
        r["n_hits"], r["dist"], r["ov"] = nh, dist, ov
        rows.append(r)
    want = _py_aggregate(rows)
    assert e.agg_n_records() == len(rows)
    assert got == want
    e.close()


def test_partition_by_key(torch_cuda):
    """hash partition for the multi-GPU exchange: every key goes to exactly one rank, stream order kept per rank"""
    torch = torch_cuda
    from find_circ2_b200._lib import JREC_DTYPE

    recs = _random_records(50000, 700, 12)
    e = _engine()
    e.agg_reset()
    e.agg_append_host(recs)
    outbuf = torch.zeros(len(recs) * 48, dtype=torch.uint8, device="cuda:0")
    for n_ranks in (1, 2, 3, 8):
        counts = e.agg_partition(n_ranks, outbuf, 0)
        torch.cuda.synchronize()
        assert counts.sum() == len(recs)
        out = outbuf.cpu().numpy().view(JREC_DTYPE)
        seen = {}
        pos = 0
        for r in range(n_ranks):
            part = out[pos : pos + counts[r]]
            pos += counts[r]
            assert (np.diff(part["idx"].astype(np.int64)) > 0).all()
            for k in set(zip(part["chrom"].tolist(), part["start"].tolist(), part["end"].tolist(), (part["sk"] & 3).tolist())):
                assert k not in seen
                seen[k] = r
        assert sorted(out["idx"].tolist()) == sorted(recs["idx"].tolist())
    e.close()


def test_tile_store_rebuild_across_read_lengths():
    """one engine, batches of different read lengths: the tile store is rebuilt for longer windows and keeps
    answering shorter ones (T = 1 -> 2 -> 4 sectors per tile)"""
    g = synth.make_genome([60000, 45000, 7000], seed=401, n_frac=0.01, n_run=(20, 200))
    J = synth.plant_junctions(g, 80, 50, seed=402, span=(150, 8000), margin=300)
    e = _engine(asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    gs = H.GenomeStrings(g)
    opt = O.Options(asize=20)
    for read_len in (76, 100, 150, 100, 250, 60, 300):
        t = synth.make_pairs(g, J, 1500, read_len=read_len, asize=20, seed=400 + read_len, error_rate=0.01,
                             frac_edge=0.05, frac_read_n=0.03)
        chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
        want = H.oracle_scan(gs, g.names, chrom, a_start, b_end, l, flags, internal, opt)
        hits = e.scan_host(chrom, a_start, b_end, l, flags, internal)
        got = [H.decode_hit(r) for r in hits.view(np.uint32).reshape(-1, 4)]
        assert got == want, read_len
    e.close()


@pytest.mark.parametrize("cap", [None, "64"])
@pytest.mark.parametrize("n,n_keys,seed,known", [(37, 5, 2, False), (200000, 5000, 4, False), (300000, 3, 5, False), (250000, 40000, 6, True)])
def test_partitioned_distinct_counts_match_python(monkeypatch, n, n_keys, seed, known, cap):
    """large inputs count distinct reads / names through partitions + shared-memory sets (agg.cu: distinct_parts_kernel)
    instead of one global set: forced here at small sizes; cap=64 makes the partitions overflow, so the call has to find
    its way back to the global set; `known`: half of the records carry the scan kernel's own verdict on their name"""
    monkeypatch.setenv("FC_AGG_SETS", "part")
    if cap:
        monkeypatch.setenv("FC_AGG_PART_CAP", cap)
    recs = _random_records(n, n_keys, seed, dens=(1, 1, 2, 4, 8))
    if known:
        # FC_SK_NAME_KNOWN: the emitter says whether the fragment is new to the junction (first record of a (junction, name))
        # every second record settles its name itself, so its name must not meet the others': make the two kinds disjoint
        recs["qname_hash"][0::2] |= np.uint64(1 << 63)
        recs["qname_hash"][1::2] &= np.uint64((1 << 63) - 1)
        first = set()
        flags = np.zeros(n, dtype=np.uint32)
        for i in range(0, n, 2):
            k = (int(recs["chrom"][i]), int(recs["start"][i]), int(recs["end"][i]), int(recs["sk"][i]) & 3, int(recs["qname_hash"][i]))
            flags[i] = 8 if k not in first else 8 | 16
            first.add(k)
        recs["sk"] |= flags
    e = _engine()
    for _ in range(2):  # twice: the partition buffers must come back clean
        e.agg_reset()
        e.agg_append_host(recs)
        nj = e.agg_finalize()
        got = _junction_rows(e.agg_fetch(nj))
        want = _py_aggregate(recs)
        assert len(got) == len(want)
        for gr, wr in zip(got, want):
            assert gr == wr, (gr, wr)
    e.close()


def test_sort_based_and_hash_based_aggregation_agree():
    """fc_agg_finalize has two implementations (sort-free default, sort-based fallback for weight denominators that are
    not powers of two): both must produce identical junction tables"""
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
from test_gpu_parity import _random_records, _engine
recs = _random_records(250000, 4000, 77, dens=(1, 1, 2, 4, 8))
e = _engine(); e.agg_reset(); e.agg_append_host(recs)
nj = e.agg_finalize(); j = e.agg_fetch(nj)
sys.stdout.buffer.write(j.tobytes())
'''
    outs = []
    for mode in ("hash", "sort"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, env=dict(os.environ, FC_AGG_MODE=mode), cwd=H.ROOT)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        outs.append(r.stdout)
    assert len(outs[0]) > 64 * 1000
    assert outs[0] == outs[1]


def _emit_records(torch, e, recs, idx_base, rng, batches=3, reject_frac=0.2):
    """feed fc_jrec rows through fc_agg_emit (hits + payload columns on the device), interleaved with pairs that found no
    breakpoint; idx of the records must be idx_base + row position in the interleaved stream -> returns the rows with
    their idx rewritten accordingly"""
    from find_circ2_b200._lib import HIT_DTYPE

    n_acc = len(recs)
    n = n_acc + int(n_acc * reject_frac) + 3
    accept = np.zeros(n, dtype=bool)
    accept[rng.choice(n, size=n_acc, replace=False)] = True
    pos = np.flatnonzero(accept)
    hits = np.zeros(n, dtype=HIT_DTYPE)
    hits["start"][pos] = recs["start"].astype(np.int64).astype(np.int32)
    hits["end"][pos] = recs["end"].astype(np.int64).astype(np.int32)
    sk = recs["sk"].astype(np.uint32)
    hits["w2"][pos] = recs["n_hits"].astype(np.uint32) | (recs["dist"].astype(np.uint32) << 16) | (recs["ov"].astype(np.uint32) << 24)
    hits["w3"][pos] = (sk & 1) | (((sk >> 16) & 0xFFF) << 1)
    back = (sk & 2) == 0
    chrom = np.zeros(n, dtype=np.int32)
    chrom[pos] = recs["chrom"]
    flags = np.zeros(n, dtype=np.uint8)
    flags[pos] = back.astype(np.uint8)  # FC_PF_BACKSPLICE = 1
    wden = np.ones(n, dtype=np.uint8)
    wden[pos] = (sk >> 8) & 0xFF
    q_a = np.zeros(n, dtype=np.int16)
    q_b = np.zeros(n, dtype=np.int16)
    q_a[pos] = np.where(back, recs["q_right"], recs["q_left"])
    q_b[pos] = np.where(back, recs["q_left"], recs["q_right"])
    rh = rng.integers(0, 1 << 62, n).astype(np.uint64)
    qh = rng.integers(0, 1 << 62, n).astype(np.uint64)
    rh[pos] = recs["read_hash"]
    qh[pos] = recs["qname_hash"]
    dev = torch.device("cuda:0")
    tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_hits = tn(hits.view(np.int32))
    cols = [tn(chrom), tn(flags), tn(wden), tn(q_a), tn(q_b), tn(rh.view(np.int64)), tn(qh.view(np.int64))]
    cuts = np.linspace(0, n, batches + 1).astype(int)
    for lo, hi in zip(cuts, cuts[1:]):
        if hi > lo:
            e.agg_emit(hi - lo, d_hits[4 * lo:], *[c[lo:] for c in cols], idx_base + lo, 0)
    out = recs.copy()
    out["idx"] = idx_base + pos
    # bit 2 of sk mirrors the palindrome flag of the read hash (set by the emit kernel)
    out["sk"] = (out["sk"] & ~np.uint32(4)) | ((out["read_hash"] & np.uint64(1)).astype(np.uint32) << 2)
    return out


POW2 = (1, 1, 1, 2, 2, 4, 8)


@pytest.mark.parametrize("hot", [False, True])
def test_sort_free_aggregation_repeated_calls(torch_cuda, hot):
    """the sort-free path keeps its key table and accumulators between calls (the finish kernel clears what it consumed):
    many calls of different sizes on one engine, records in stream order (fc_agg_emit) and in arbitrary
    order (fc_agg_append), with and without one junction that collects a third of the records"""
    torch = torch_cuda
    e = _engine()
    rng = np.random.default_rng(99)
    for rep, (n, n_keys) in enumerate([(40000, 900), (700, 40), (150000, 30000), (5, 2), (90000, 17), (150000, 2500), (1, 1),
                                       (60000, 60000)]):
        recs = _random_records(n, n_keys, 1000 + rep, dens=POW2)
        if hot and n > 100:
            sel = rng.random(n) < 0.34
            for f in ("chrom", "start", "end"):
                recs[f][sel] = recs[f][0]
            recs["sk"][sel] = (recs["sk"][sel] & ~np.uint32(3)) | (recs["sk"][0] & np.uint32(3))
        e.agg_reset()
        if rep % 2 == 0:
            recs = _emit_records(torch, e, recs, 5000 + rep, rng)
        else:
            order = rng.permutation(n)
            e.agg_append_host(recs[order])
            if rep % 4 == 1:
                # the caller knows the idx range of what it appended: discovery rank from flags instead of a sort
                e.agg_set_idx_range(int(recs["idx"].min()), int(recs["idx"].max()) + 1)
        nj = e.agg_finalize()
        got = _junction_rows(e.agg_fetch(nj))
        want = _py_aggregate(recs[np.argsort(recs["idx"], kind="stable")])
        assert e.agg_n_records() == n
        assert len(got) == len(want), (rep, n, len(got), len(want))
        assert got == want, rep
        assert e.agg_finalize() == nj  # idempotent: the tables were left clean
        assert _junction_rows(e.agg_fetch(nj)) == want
    e.close()


def test_declared_idx_range_that_is_too_small_still_gives_the_right_table(torch_cuda):
    """fc_agg_set_idx_range is a promise of the caller; a record outside it must not corrupt anything: the call falls back
    to ranking by sort and the next calls on the engine work as usual"""
    e = _engine()
    rng = np.random.default_rng(5)
    recs = _random_records(30000, 500, 77, dens=POW2)
    want = _py_aggregate(recs[np.argsort(recs["idx"], kind="stable")])
    lo, hi = int(recs["idx"].min()), int(recs["idx"].max()) + 1
    for declared in ((lo + 1000, hi), (lo, lo + (hi - lo) // 2), (lo, hi)):
        e.agg_reset()
        e.agg_append_host(recs[rng.permutation(len(recs))])
        e.agg_set_idx_range(*declared)
        nj = e.agg_finalize()
        assert _junction_rows(e.agg_fetch(nj)) == want, declared
    e.close()


@pytest.mark.parametrize("read_len,noncanonical", [(100, False), (76, False), (150, False), (250, False), (100, True)])
def test_scan_emit_equals_scan_then_emit(torch_cuda, read_len, noncanonical):
    """fc_scan_emit (one kernel) == fc_scan + fc_agg_emit: same hits, same junction table -- for every specialisation of
    the scan kernel (tile classes T=1/2/4, the per-base path of --non-canonical)"""
    torch = torch_cuda
    g, J, t = _case(read_len, 20, seed=41 + read_len, n=7001, error_rate=0.01)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    n = len(chrom)
    e = _engine(asize=20, noncanonical=noncanonical)
    e.load_genome_arrays(g.names, g.seqs)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(9)
    wden = rng.choice(np.array([1, 1, 2, 4, 8], dtype=np.uint8), size=n)
    q_a = (t.as_a - np.maximum(t.xs_a, 0)).astype(np.int16)
    q_b = (t.as_b - np.maximum(t.xs_b, 0)).astype(np.int16)
    rh = e.hash_reads(t.reads, np.full(n, t.read_len, dtype=np.int32))
    qh = np.array([e.hash_bytes(("q%d" % (i // 3)).encode()) for i in range(n)], dtype=np.uint64)
    n_words = max(1, (int(l.max()) + 31) // 32)
    tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_chrom, d_a, d_b, d_l, d_fl, d_asc = tn(chrom), tn(a_start), tn(b_end), tn(l), tn(flags), tn(internal)
    planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    e.pack_reads(d_asc, internal.shape[1], d_l, n_words, planes, d_fl, 0)
    pairs = e.make_pairs(n, d_chrom, d_a, d_b, d_l, d_fl, planes, n_words, int(l.max()))
    pay = (tn(wden), tn(q_a), tn(q_b), tn(rh.view(np.int64)), tn(qh.view(np.int64)))
    out1 = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    out2 = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    e.agg_reset()
    e.scan(pairs, out1, 0)
    e.agg_emit(n, out1, d_chrom, d_fl, *pay, 500, 0)
    t1 = e.agg_fetch(e.agg_finalize())
    n1 = e.agg_n_records()
    e.agg_reset()
    e.scan_emit(pairs, out2, *pay, 500, 0)
    t2 = e.agg_fetch(e.agg_finalize())
    assert e.agg_n_records() == n1 and n1 > 1000
    assert torch.equal(out1, out2)
    assert len(t1) > 50 and t1.tobytes() == t2.tobytes()
    e.close()


@pytest.mark.parametrize("read_len", [100, 150, 44])
def test_packed_batch_with_fragment_rows_equals_soa_path(torch_cuda, read_len):
    """fc_batch_pack + fc_scan_emit_batch with the fragment fields of the descriptors (the scan kernel settles n_frags with
    shuffles, records carry FC_SK_NAME_KNOWN) == fc_scan_emit on the fc_pairs batch (every record goes through the name set):
    same hits, same junction table.  Fragments of 1..6 rows, rows of one fragment often on the same junction, fragments
    straddling warp boundaries (32 rows) and one that is too long for the 2-bit fields."""
    torch = torch_cuda
    g, J, t = _case(read_len, 20, seed=77 + read_len, n=9000, error_rate=0.01)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    n = len(chrom)
    rng = np.random.default_rng(5)
    # fragments: runs of rows that share a read name; make the rows of a run copies of its first row now and then
    sizes = rng.choice([1, 1, 1, 2, 2, 3, 4, 5, 6], size=n)
    frag_id = np.repeat(np.arange(len(sizes)), sizes)[:n]
    first = np.concatenate([[0], np.nonzero(np.diff(frag_id))[0] + 1])
    first_of = first[np.searchsorted(first, np.arange(n), side="right") - 1]
    copy = rng.random(n) < 0.5
    src = np.where(copy, first_of, np.arange(n))
    chrom, a_start, b_end, l, flags, internal = (x[src] for x in (chrom, a_start, b_end, l, flags, internal))
    reads = t.reads[src]
    e = _engine(asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    dev = torch.device("cuda:0")
    q_a = (t.as_a - np.maximum(t.xs_a, 0)).astype(np.int16)[src]
    q_b = (t.as_b - np.maximum(t.xs_b, 0)).astype(np.int16)[src]
    wden = rng.choice(np.array([1, 1, 2, 4], dtype=np.uint8), size=n)
    rh = e.hash_reads(reads, np.full(n, t.read_len, dtype=np.int32))
    qh = np.array([e.hash_bytes(b"frag%d" % f) for f in frag_id], dtype=np.uint64)
    back = np.arange(n) - first_of
    size_of = np.bincount(frag_id)[frag_id]
    fwd = size_of - 1 - back
    frag = np.where(size_of > 4, 15, back | (fwd << 2)).astype(np.uint8)
    n_words = max(1, (int(l.max()) + 31) // 32)
    tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_chrom, d_a, d_b, d_l, d_fl, d_asc = tn(chrom), tn(a_start), tn(b_end), tn(l), tn(flags), tn(internal)
    planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    e.pack_reads(d_asc, internal.shape[1], d_l, n_words, planes, d_fl, 0)
    pairs = e.make_pairs(n, d_chrom, d_a, d_b, d_l, d_fl, planes, n_words, int(l.max()))
    pay = (tn(wden), tn(q_a), tn(q_b), tn(rh.view(np.int64)), tn(qh.view(np.int64)))
    out1 = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    out2 = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    e.agg_reset()
    e.scan_emit(pairs, out1, *pay, 0, 0)
    t1 = e.agg_fetch(e.agg_finalize())
    nw = e.batch_words(int(l.max()))
    meta = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    rows = torch.zeros(n * 2 * nw, dtype=torch.int32, device=dev)
    rn_rows = torch.zeros(n * nw, dtype=torch.int32, device=dev)
    q = torch.zeros(n, dtype=torch.int32, device=dev)
    for fr in (None, tn(frag)):
        batch = e.pack_batch(pairs, meta, rows, rn_rows, pay[0], pay[1], pay[2], q, fr, 0)
        e.agg_reset()
        e.scan_emit_batch(batch, out2, q, pay[3], pay[4], 0, 0)
        t2 = e.agg_fetch(e.agg_finalize())
        assert torch.equal(out1, out2)
        assert len(t1) > 50 and t1.tobytes() == t2.tobytes()
    assert (t1["n_frags"] < t1["n_spanned"]).any()
    out3 = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    e.scan_batch(batch, out3, 0)
    assert torch.equal(out1, out3)
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sets", ["global", "part"])
def test_junction_table_sized_after_the_last_batch(monkeypatch, sets):
    """a context that has reduced a batch sizes its junction table after that batch's junction count (agg.cu: nj_hint); a
    batch with many more junctions than the table holds must be noticed and reduced again with the full table -- and the
    calls after it must find clean tables"""
    monkeypatch.setenv("FC_AGG_SETS", sets)
    e = _engine()
    small = _random_records(20000, 50, 21, dens=(1, 2, 4, 8))
    big = _random_records(300000, 120000, 22, dens=(1, 2, 4, 8))
    for recs in (small, small, big, big, small):
        e.agg_reset()
        e.agg_append_host(recs)
        nj = e.agg_finalize()
        got = _junction_rows(e.agg_fetch(nj))
        want = _py_aggregate(recs)
        assert len(got) == len(want)
        for gr, wr in zip(got, want):
            assert gr == wr, (gr, wr)
    # a table far too small on purpose (every record runs out of probes at once)
    monkeypatch.setenv("FC_AGG_TABLE_HINT", "1024")
    for recs in (big, small):
        e.agg_reset()
        e.agg_append_host(recs)
        nj = e.agg_finalize()
        assert _junction_rows(e.agg_fetch(nj)) == _py_aggregate(recs)
    e.close()
