"""
BASELINE.json full-size checks on the GPU through size-independent properties (the oracle needs ~0.3 ms per pair, so it
only checks a random sample here):
  config 2   100 Mb genome, 10 000 planted circRNAs, 1 M anchor pairs, 100-nt reads
  config 3*  hg19-sized coordinates: a 3.1 Gb genome (global bit indices beyond 2^31), 2 M pairs, 2x100-nt style reads
"""
import numpy as np
import pytest

import helpers as H
from find_circ2_b200 import synth
from oracle import find_circ_oracle as O

pytestmark = pytest.mark.gpu

ASIZE, MARGIN = 20, 2


def _soa(t):
    return H.pairs_to_soa(t, ASIZE, MARGIN)


def _intact(g, J):
    """junctions whose planted dinucleotides survived later plantings (two junctions a base apart overwrite each other)"""
    ok = np.zeros(len(J), dtype=bool)
    for i in range(len(J)):
        seq = g.seqs[int(J.chrom[i])]
        s, e = int(J.start[i]), int(J.end[i])
        if J.circ[i]:
            left, right = (b"AC", b"CT") if J.minus[i] else (b"AG", b"GT")
            ok[i] = seq[s - 2 : s].tobytes() == left and seq[e : e + 2].tobytes() == right
        else:
            left, right = (b"CT", b"AC") if J.minus[i] else (b"GT", b"AG")
            ok[i] = seq[s : s + 2].tobytes() == left and seq[e - 2 : e].tobytes() == right
    return ok


def _check_planted(g, J, t, hits, chrom, flags):
    """a read of a planted junction that carries no sequencing error and no N must recover exactly that junction"""
    eff = ASIZE - MARGIN
    R = t.read_len
    intact = _intact(g, J)
    ok = (t.junc >= 0) & intact[np.maximum(t.junc, 0)]
    # error-free: the read equals the genome on both sides -> dist 0.  We only know that for reads we did not mutate;
    # use dist==0 hits and require coordinates == planted coordinates whenever the true split is inside the scan range.
    nh = hits["w2"] & 0xFFFF
    dist = (hits["w2"] >> 16) & 0xFF
    sel = np.nonzero(ok)[0]
    good = sel[(nh[sel] == 1) & (dist[sel] == 0)]
    assert len(good) > 0.5 * len(sel)
    jj = t.junc[good]
    same = ((hits["start"][good].astype(np.int64) == J.start[jj]) & (hits["end"][good].astype(np.int64) == J.end[jj])
            & ((hits["w3"][good] & 1).astype(bool) == J.minus[jj]) & ((flags[good] & 1).astype(bool) == J.circ[jj]))
    # the rare exceptions (a second planted junction inside the read, chance homology next to a damaged signal) must be
    # exceptions for the oracle too
    odd = good[~same]
    assert len(odd) < 2e-3 * len(good), (len(odd), len(good))
    return odd[:300]


def _full_case(engine_kw, genome_sizes, n_pairs, n_circ, seed, sample=1500):
    from find_circ2_b200.engine import Engine

    g = synth.make_genome(genome_sizes, seed=seed, n_frac=0.005, n_run=(50, 5000), soft_frac=0.0)
    J = synth.plant_junctions(g, n_circ, n_circ // 20, seed=seed + 1, span=(200, 50000), margin=400)
    t = synth.make_pairs(g, J, n_pairs, read_len=100, asize=ASIZE, seed=seed + 2, error_rate=0.005, frac_decoy=0.10,
                         frac_nonuniq=0.0, frac_edge=0.01)
    chrom, a_start, b_end, l, flags, internal = _soa(t)
    e = Engine(device=0, asize=ASIZE, **engine_kw)
    e.load_genome_arrays(g.names, g.seqs)
    n = len(chrom)
    rh = e.hash_reads(t.reads, np.full(n, t.read_len, dtype=np.int32))
    qh = (t.name_id.astype(np.uint64) // np.uint64(2)) * np.uint64(0x9E3779B97F4A7C15)  # two reads share a fragment name
    qa = (t.as_a - np.maximum(t.xs_a, 0)).astype(np.int16)
    qb = (t.as_b - np.maximum(t.xs_b, 0)).astype(np.int16)
    e.agg_reset()
    hits = e.batch_host(chrom, a_start, b_end, l, flags, internal, np.ones(n, np.uint8), qa, qb, rh, qh, 0, emit=True)
    # idempotence: the same batch scanned again gives the same hits
    again = e.scan_host(chrom, a_start, b_end, l, flags, internal)
    assert np.array_equal(hits.view(np.uint32), again.view(np.uint32))
    odd = _check_planted(g, J, t, hits, chrom, flags)
    # a random sample (plus every pair that did not recover its planted junction) against the oracle
    rng = np.random.default_rng(seed)
    idx = np.unique(np.concatenate([rng.choice(n, size=sample, replace=False), odd]))
    want = H.oracle_scan(H.GenomeStrings(g), g.names, chrom[idx], a_start[idx], b_end[idx], l[idx], flags[idx], internal[idx],
                         O.Options(asize=ASIZE))
    got = [H.decode_hit(r) for r in hits[idx].view(np.uint32).reshape(-1, 4)]
    assert got == want
    # aggregation: counts per key equal a numpy group-by of the hits; distinct counts are bounded by them
    nh = (hits["w2"] & 0xFFFF) > 0
    nj = e.agg_finalize()
    junc = e.agg_fetch(nj)
    assert e.agg_n_records() == int(nh.sum())
    assert int(junc["n_spanned"].sum()) == int(nh.sum())
    key = np.stack([chrom[nh].astype(np.int64), hits["start"][nh].astype(np.int64), hits["end"][nh].astype(np.int64),
                    (hits["w3"][nh] & 1).astype(np.int64), (1 - (flags[nh] & 1)).astype(np.int64)], axis=1)
    uk, first, cnt = np.unique(key, axis=0, return_index=True, return_counts=True)
    assert nj == len(uk)
    got_key = np.stack([junc["chrom"].astype(np.int64), junc["start"].astype(np.int64), junc["end"].astype(np.int64),
                        (junc["sk"] & 1).astype(np.int64), ((junc["sk"] >> 1) & 1).astype(np.int64)], axis=1)
    order = np.lexsort(got_key.T[::-1])
    assert np.array_equal(got_key[order], uk)
    assert np.array_equal(junc["n_spanned"][order].astype(np.int64), cnt)
    # discovery order: first_idx is the row of the first supporting pair, rows come sorted by it
    rows = np.nonzero(nh)[0]
    assert np.array_equal(junc["first_idx"][order].astype(np.int64), rows[first])
    assert (np.diff(junc["first_idx"].astype(np.int64)) > 0).all()
    # n_uniq may legitimately be 0: junctions planted inside an N run are supported by all-N reads, which equal their
    # own reverse complement -- {read, rev_comp(read)} has ONE element and the reference prints len/2 = 0 (find_circ.py:588-590)
    assert (junc["n_uniq"] <= junc["n_spanned"]).all()
    assert (junc["n_frags"] <= junc["n_spanned"]).all() and (junc["n_frags"] >= 1).all()
    assert np.array_equal(junc["n_weighted"], junc["n_spanned"].astype(np.float64))
    # exact distinct counts for the 50 biggest junctions
    big = np.argsort(-junc["n_spanned"].astype(np.int64))[:50]
    inv = {tuple(k): i for i, k in enumerate(got_key.tolist())}
    where = np.array([inv[tuple(k)] for k in key.tolist()])
    rhh, qhh = rh[nh], qh[nh]
    for j in big:
        m = where == j
        assert int(junc["n_frags"][j]) == len(np.unique(qhh[m]))
        pal = np.unique(rhh[m][(rhh[m] & np.uint64(1)) == 1])
        assert int(junc["n_uniq"][j]) == len(np.unique(rhh[m])) - (len(pal) + 1) // 2
    # the same batch through the chunked host-buffer call that takes bit-plane reads (two-stream chunks, scan kernel that
    # emits on the way): same hits, same junction table
    import torch

    dev = torch.device("cuda:0")
    max_l = int(l.max())
    n_words = max(1, (max_l + 31) // 32)
    d_planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    d_fl = torch.from_numpy(flags.copy()).to(dev)
    e.pack_reads(torch.from_numpy(np.ascontiguousarray(internal)).to(dev), internal.shape[1], torch.from_numpy(l).to(dev), n_words,
                 d_planes, d_fl, 0)
    torch.cuda.synchronize()
    planes = d_planes.cpu().numpy().view(np.uint32).reshape(3, n_words * n)
    e.agg_reset()
    hits2 = e.batch_host_planes(n, chrom, a_start, b_end, l, d_fl.cpu().numpy(), planes[0], planes[1], planes[2], n_words, n, max_l,
                                np.ones(n, np.uint8), qa, qb, rh, qh, idx=None, idx_base=0, emit=True)
    assert np.array_equal(hits.view(np.uint32), hits2.view(np.uint32))
    junc2 = e.agg_fetch(e.agg_finalize())
    assert junc2.tobytes() == junc.tobytes()
    st = e.genome_stats()
    e.close()
    return st


def test_config2_full_size():
    st = _full_case({}, [5000000] * 20, 1000000, 10000, seed=1)
    assert st["bases"] == 100000000


def test_hg19_sized_coordinates():
    """3.1 Gb of genome: global base indices pass 2^31; tile store ~3.2 GB + planes ~1.2 GB on the device"""
    import psutil

    if psutil.virtual_memory().available < 24 << 30:
        pytest.skip("needs ~20 GB of host memory to synthesise the genome")
    sizes = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022, 141213431, 135534747,
             135006516, 133851895, 115169878, 107349540, 102531392, 90354753, 81195210, 78077248, 59128983, 63025520,
             48129895, 51304566, 155270560]  # chr1..22, X (test_data/test_norm.sam:1-93 lists the hg19 lengths)
    st = _full_case({}, sizes, 2000000, 100000, seed=5, sample=800)
    assert st["bases"] == sum(sizes)


def test_config5_shaped_aggregation(monkeypatch):
    """BASELINE.json configs[4] in miniature on one GPU: 6 M junction records over 1.2 M distinct junctions with a Zipf(0.6)
    tail (most junctions carry 1-5 reads), repeated read sequences and fragment names, weights 1, 1/2, 1/4, 1/8 -- the
    one-pass table (64-byte slots, exact hash set) and the sort-based reduce must give the same table, byte for byte, and the
    table must equal a numpy group-by"""
    from find_circ2_b200._lib import JREC_DTYPE
    from find_circ2_b200.engine import Engine

    rng = np.random.default_rng(55)
    n, nj = 6000000, 1200000
    w = 1.0 / np.power(np.arange(1, nj + 1, dtype=np.float64), 0.6)
    jx = rng.choice(nj, size=n, p=w / w.sum())
    jc = rng.integers(0, 24, size=nj).astype(np.uint32)
    js = rng.integers(1000, 200000000, size=nj).astype(np.int32)
    je = (js + rng.integers(200, 50000, size=nj)).astype(np.int32)
    jk = rng.integers(0, 4, size=nj).astype(np.uint32)
    recs = np.zeros(n, dtype=JREC_DTYPE)
    recs["chrom"], recs["start"], recs["end"] = jc[jx], js[jx], je[jx]
    den = rng.choice(np.array([1, 1, 1, 2, 4, 8], dtype=np.uint32), size=n)
    rh = rng.integers(0, 1 << 62, size=n, dtype=np.int64).astype(np.uint64) << np.uint64(1)
    dup = rng.random(n) < 0.1
    rh[dup] = rh[rng.integers(0, n, size=int(dup.sum()))]  # repeated reads (some inside one junction, most not)
    rh[rng.random(n) < 0.01] |= np.uint64(1)                # palindromic reads
    recs["sk"] = jk[jx] | ((rh & np.uint64(1)).astype(np.uint32) << 2) | (den << 8) | (np.uint32(0x2D2) << 16)
    recs["idx"] = np.arange(n, dtype=np.uint64) + np.uint64(1000)
    recs["read_hash"] = rh
    recs["qname_hash"] = (np.arange(n, dtype=np.uint64) // np.uint64(2)) * np.uint64(0x9E3779B97F4A7C15)
    recs["q_left"] = rng.integers(0, 60, size=n)
    recs["q_right"] = rng.integers(0, 60, size=n)
    recs["n_hits"] = rng.integers(1, 4, size=n)
    recs["dist"] = rng.integers(0, 3, size=n)
    recs["ov"] = rng.integers(0, 3, size=n)
    e = Engine(device=0, asize=20)
    tables = []
    for mode in ("", "sort"):
        monkeypatch.setenv("FC_AGG_MODE", mode)
        e.agg_reset()
        e.agg_append_host(recs)
        tables.append(e.agg_fetch(e.agg_finalize()))
    monkeypatch.setenv("FC_AGG_MODE", "")
    a, b = tables
    assert len(a) == len(np.unique(jx)) and a.tobytes() == b.tobytes()
    # against numpy: counts, weights and distinct reads / names of every junction
    key = jx
    order = np.argsort(key, kind="stable")
    ks = key[order]
    starts = np.concatenate([[0], np.nonzero(np.diff(ks))[0] + 1])
    cnt = np.diff(np.concatenate([starts, [n]]))
    first = recs["idx"][order][starts]
    by_first = np.argsort(first)
    assert np.array_equal(a["first_idx"], first[by_first])
    assert np.array_equal(a["n_spanned"].astype(np.int64), cnt[by_first])
    wsum = np.add.reduceat((1.0 / den[order]), starts)
    assert np.array_equal(a["n_weighted"], wsum[by_first])
    pair = np.stack([ks, rh[order].view(np.int64)], axis=1)
    up = np.unique(pair, axis=0)
    n_reads = np.bincount(np.searchsorted(np.unique(ks), up[:, 0]), minlength=len(starts))
    pal = up[(up[:, 1] & 1) == 1]
    n_pal = np.bincount(np.searchsorted(np.unique(ks), pal[:, 0]), minlength=len(starts))
    assert np.array_equal(a["n_uniq"].astype(np.int64), (n_reads - (n_pal + 1) // 2)[by_first])
    qn = np.stack([ks, recs["qname_hash"][order].view(np.int64)], axis=1)
    n_names = np.bincount(np.searchsorted(np.unique(ks), np.unique(qn, axis=0)[:, 0]), minlength=len(starts))
    assert np.array_equal(a["n_frags"].astype(np.int64), n_names[by_first])
    e.close()
