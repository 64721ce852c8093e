"""fc_ingest_evidence (C++, csrc/ingest.cu) against its numpy twin (tests/evidence_numpy.py) on random fragment records:
every combination of queued / back-splice / hit / unspliced-mate / broken flags that a two-span fragment can show.
The reference's own answers for these rules are pinned by the goldens (test_pipeline_host_logic.py, test_gpu_pipeline.py)."""
import numpy as np
import pytest

from evidence_numpy import evidence_numpy
from find_circ2_b200 import _lib
from find_circ2_b200._lib import HIT_DTYPE
from find_circ2_b200.ingest import evidence
from find_circ2_b200.pipeline import FLAG_BIT


def _random_batch(m, seed):
    rng = np.random.default_rng(seed)
    state = rng.choice(np.array([0, 1, 2, 3], dtype=np.uint8), m, p=[0.05, 0.45, 0.1, 0.4])
    nrows = (state & 1) + ((state >> 1) & 1)
    row0 = np.zeros(m, dtype=np.int32)
    row0[1:] = np.cumsum(nrows[:-1])
    n = int(nrows.sum())
    a = {"f_state": state, "f_row0": row0, "f_nsp": rng.integers(1, 3, m).astype(np.uint8), "f_kind": rng.integers(0, 4, m).astype(np.uint8),
         "f_flags": rng.integers(0, 16, m).astype(np.uint8), "f_un_tid": rng.integers(0, 3, m).astype(np.int32),
         "f_un_pos": rng.integers(0, 3000, m).astype(np.int32), "f_seq": np.arange(m, dtype=np.int64) * 3 + 5,
         "f_txt_off": rng.integers(0, 1 << 20, 6 * m).astype(np.int64), "f_txt_len": rng.integers(-1, 120, 6 * m).astype(np.int32),
         "chrom": rng.integers(0, 3, max(n, 1)).astype(np.int32), "qname_hash": rng.integers(0, 1 << 62, max(n, 1)).astype(np.uint64)}
    a["f_un_aend"] = (a["f_un_pos"] + rng.integers(20, 200, m)).astype(np.int32)
    hits = np.zeros(max(n, 1), dtype=HIT_DTYPE)
    # few distinct coordinates: the two spans of a fragment often name the same junction, nest or lie apart
    hits["start"] = rng.integers(0, 8, len(hits)) * 400
    hits["end"] = hits["start"] + rng.integers(1, 6, len(hits)) * 400
    hits["w2"] = np.where(rng.random(len(hits)) < 0.8, rng.integers(1, 4, len(hits)), 0) | (rng.integers(0, 3, len(hits)) << 16)
    hits["w3"] = rng.integers(0, 2, len(hits)) | (0x4D3 << 1)
    return a, hits, n


@pytest.mark.parametrize("m,seed", [(1, 1), (7, 2), (1000, 3), (50000, 4)])
def test_evidence_pass_equals_numpy_twin(m, seed):
    lib = _lib.load()
    a, hits, n = _random_batch(m, seed)
    got = evidence(lib, a, hits, n, m, 1000, 20, FLAG_BIT)
    want = evidence_numpy(a, hits, n, m, 1000, 20, FLAG_BIT)
    assert got["counters"] == want["counters"] and got["any_hit"] == want["any_hit"]
    for k in ("W", "cls", "key0", "key1", "ck", "r_seq", "r_k0", "r_k1", "r_mask", "r_off3", "r_len3"):
        assert np.array_equal(got[k], want[k]), k
    # the events come in another order (the product sorts them per junction anyway)
    def canon(r):
        rows = np.concatenate([r["ev_key"], r["ev_hash"].astype(np.int64)[:, None], r["ev_mask"].astype(np.int64)[:, None]], axis=1)
        return rows[np.lexsort(rows.T[::-1])]
    assert np.array_equal(canon(got), canon(want))
    if m >= 1000:
        assert len(got["ev_mask"]) > 0 and len(got["r_seq"]) > m // 2
        assert len(set(got["W"].tolist())) > 8  # the flags really vary


def test_native_ingest_planes_for_every_kind_of_letter():
    """the ingest's word-at-a-time packer (csrc/ingest.cu: pack8) against the plain definition of the planes: bit j of word w =
    code of base 32 w + j of the internal read part (A 0, C 1, G 2, T 3, lower case alike; anything else: N plane, code 0)"""
    from find_circ2_b200.ingest import NativeIngest

    rng = np.random.default_rng(11)
    letters = np.frombuffer(b"ACGTacgtNnRYKMSWBDHVXacgtACGTACGTACGTACGTACGT", dtype=np.uint8)
    asize, margin, L = 20, 2, 150
    eff = asize - margin
    lines, reads = [], []
    for k in range(300):
        seq = bytes(rng.choice(letters, L)).decode()
        cut = int(rng.integers(asize, L - asize))
        pos_a, pos_b = 5000 + 3 * k, 1000 + 3 * k
        lines.append("r%d\t0\tchr1\t%d\t60\t%dM%dS\t*\t0\t0\t%s\t*\tAS:i:%d\tXS:i:0\n" % (k, pos_a + 1, cut, L - cut, seq, cut))
        lines.append("r%d\t2048\tchr1\t%d\t60\t%dH%dM\t*\t0\t0\t%s\t*\tAS:i:%d\tXS:i:0\n" % (k, pos_b + 1, cut, L - cut, seq[cut:], L - cut))
        reads.append(seq)
    text = ("@SQ\tSN:chr1\tLN:100000\n" + "".join(lines)).encode()
    ing = NativeIngest(asize, margin, 2, False, ["chr1"], [0], cap=1024, n_words=8)
    for rep in range(2):  # twice into the same arrays: what the first call wrote must not shine through
        used = ing.parse(text, 0, True)
        assert used == len(text)
        n = int(ing.out.n_rows)
        assert n == 300 and int(ing.out.n_complex) == 0
        planes = {k: ing.a[k].reshape(8, ing.cap) for k in ("rlo", "rhi", "rn")}
        for i, seq in enumerate(reads):
            internal = seq[eff:L - eff]
            assert int(ing.a["l"][i]) == len(internal)
            for w in range(8):
                lo = hi = nn = 0
                for j, ch in enumerate(internal[32 * w:32 * w + 32]):
                    code = "ACGT".find(ch.upper())
                    if code < 0:
                        nn |= 1 << j
                        code = 0
                    lo |= (code & 1) << j
                    hi |= (code >> 1) << j
                assert (int(planes["rlo"][w, i]), int(planes["rhi"][w, i]), int(planes["rn"][w, i])) == (lo, hi, nn), (i, w)
        # second round: shorter reads in the same rows (fewer words per row)
        reads = [s[:100] for s in reads]
        L2 = 100
        lines = []
        for k, seq in enumerate(reads):
            cut = 50
            lines.append("r%d\t0\tchr1\t%d\t60\t%dM%dS\t*\t0\t0\t%s\t*\tAS:i:%d\tXS:i:0\n" % (k, 5001 + 3 * k, cut, L2 - cut, seq, cut))
            lines.append("r%d\t2048\tchr1\t%d\t60\t%dH%dM\t*\t0\t0\t%s\t*\tAS:i:%d\tXS:i:0\n" % (k, 1001 + 3 * k, cut, L2 - cut, seq[cut:], L2 - cut))
        text = ("@SQ\tSN:chr1\tLN:100000\n" + "".join(lines)).encode()
        L = L2
    ing.close()


@pytest.mark.parametrize("n,w,vals", [(0, 5, 3), (1, 5, 3), (1000, 11, 2), (200000, 5, 40), (50000, 3, 1 << 40)])
def test_unique_rows_equals_numpy(n, w, vals):
    """fc_unique_rows (C++ hash table) against np.unique(axis=0): same distinct rows, consistent inverse, first-appearance order"""
    from find_circ2_b200.pipeline import _unique_rows

    rng = np.random.default_rng(n + w)
    key = rng.integers(-2, vals, (n, w)).astype(np.int64)
    uniq, inv = _unique_rows(key)
    if n == 0:
        assert len(uniq) == 0 and len(inv) == 0
        return
    want = np.unique(key, axis=0)
    assert len(uniq) == len(want)
    assert np.array_equal(uniq[np.lexsort(uniq.T[::-1])], want)
    assert np.array_equal(uniq[inv], key)
    # order of first appearance
    firsts = [int(np.nonzero(inv == u)[0][0]) for u in range(min(len(uniq), 50))]
    assert firsts == sorted(firsts)


def test_read_hash_spreads_over_64_bits():
    """n_uniq / n_frags compare 64-bit hashes of reads and names (DESIGN.md section 2 has the collision bound, which assumes a
    hash that spreads): a million distinct reads -- random ones and runs of reads that differ in one base, the shape of real
    duplicates-with-an-error -- must give a million distinct hashes, strand-invariant, with bit 0 reserved for palindromes and the
    lower bits balanced"""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    L, n = 100, 1000000
    reads = rng.integers(0, 4, (n, L)).astype(np.uint8)
    base = reads[0].copy()
    for k in range(1, 300):  # one-base neighbours of one read
        reads[k] = base
        reads[k, k % L] = (base[k % L] + 1 + k // L) % 4
    ascii_ = np.frombuffer(b"ACGT", dtype=np.uint8)[reads]
    lens = np.full(n, L, dtype=np.int32)
    h = np.zeros(n, dtype=np.uint64)
    pal = np.zeros(n, dtype=np.uint8)
    assert lib.fc_hash_reads_host(n, ascii_.ctypes.data, L, lens.ctypes.data, h.ctypes.data, pal.ctypes.data) == 0
    assert pal.sum() == 0 and not (h & np.uint64(1)).any()
    assert len(np.unique(h)) == len(np.unique(ascii_[:, :], axis=0)) == n
    # the reverse complement of a read is the same element (find_circ.py:588-590)
    comp = np.frombuffer(b"TGCA", dtype=np.uint8)[reads[:1000, ::-1]]
    h2 = np.zeros(1000, dtype=np.uint64)
    assert lib.fc_hash_reads_host(1000, np.ascontiguousarray(comp).ctypes.data, L, lens.ctypes.data, h2.ctypes.data, None) == 0
    assert np.array_equal(h2, h[:1000])
    # (the element's hash is the smaller of the hashes of its two strands, so the top bits lean towards 0 -- one bit of the 63
    # is spent on strand invariance; the tables mix the value again before they use it)
    bits = ((h[:, None] >> np.arange(1, 56, dtype=np.uint64)[None, :]) & np.uint64(1)).mean(axis=0)
    assert np.all(np.abs(bits - 0.5) < 0.005), bits
