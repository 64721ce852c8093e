"""
fc_stream (pytest -m gpu): batches in fc_batch layout streamed from pinned host memory through the slots -- copy in, scan,
record, results back, no host wait in between -- must give the hits and the junction table of the device-resident calls
(fc_scan_batch / fc_scan_emit_batch), whatever the chunking.  Replaces the reference's per-fragment loop
(find_circ.py:1535-1574) at the C ABI.
"""
import numpy as np
import pytest

import helpers as H
from find_circ2_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


def _packed_case(torch, read_len, n, seed):
    from find_circ2_b200.engine import Engine

    dev = torch.device("cuda:0")
    g = synth.make_genome([90000, 50000], seed=seed, n_frac=0.01, n_run=(20, 200))
    J = synth.plant_junctions(g, 120, 40, seed=seed + 1, span=(150, 8000), margin=300)
    t = synth.make_pairs(g, J, n, read_len=read_len, asize=20, seed=seed + 2, error_rate=0.01, frac_read_n=0.05)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    e = Engine(device=0, asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d = [tn(x) for x in (chrom, a_start, b_end, l, flags, internal)]
    n_words = e.n_words_for(int(l.max()))
    planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    e.pack_reads(d[5], internal.shape[1], d[3], n_words, planes, d[4], 0)
    pairs = e.make_pairs(n, d[0], d[1], d[2], d[3], d[4], planes, n_words, int(l.max()))
    q_a = (t.as_a - np.maximum(t.xs_a, 0)).astype(np.int16)
    q_b = (t.as_b - np.maximum(t.xs_b, 0)).astype(np.int16)
    rh = e.hash_reads(t.reads, np.full(n, read_len, dtype=np.int32))
    qh = np.array([e.hash_bytes(b"q%d" % (i // 2)) for i in range(n)], dtype=np.uint64)
    nw = e.batch_words(int(l.max()))
    meta = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    rows = torch.zeros(n * 2 * nw, dtype=torch.int32, device=dev)
    rn_rows = torch.zeros(n * nw, dtype=torch.int32, device=dev)
    q = torch.zeros(n, dtype=torch.int32, device=dev)
    batch = e.pack_batch(pairs, meta, rows, rn_rows, None, tn(q_a), tn(q_b), q, None, 0)
    torch.cuda.synchronize()
    fl = d[4].cpu().numpy()
    return dict(e=e, n=n, nw=nw, max_l=int(l.max()), batch=batch, meta=meta, rows=rows, rn_rows=rn_rows, q=q, rh=rh, qh=qh, flags=fl,
                keep=(d, planes, pairs))


@pytest.mark.parametrize("read_len,chunk", [(100, 4000), (100, 1111), (150, 2500)])
def test_stream_equals_device_calls(torch_cuda, read_len, chunk):
    torch = torch_cuda
    from find_circ2_b200._lib import HIT_DTYPE
    from find_circ2_b200.engine import HostStream

    c = _packed_case(torch, read_len, 9000, seed=11 + read_len)
    e, n, nw = c["e"], c["n"], c["nw"]
    dev = torch.device("cuda:0")
    # device-resident answer
    hits = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    e.agg_reset()
    e.scan_emit_batch(c["batch"], hits, c["q"], torch.from_numpy(c["rh"].view(np.int64)).to(dev), torch.from_numpy(c["qh"].view(np.int64)).to(dev), 77, 0)
    want_table = e.agg_fetch(e.agg_finalize())
    want_hits = hits.cpu().numpy().view(np.uint32).reshape(n, 4)
    assert len(want_table) > 50 and (c["flags"] & 4).any()
    # host copies of the packed batch (pageable memory works too, the copies then simply do not overlap)
    h_meta, h_rows, h_q = c["meta"].cpu().numpy(), c["rows"].cpu().numpy(), c["q"].cpu().numpy()
    h_rn = c["rn_rows"].cpu().numpy().reshape(n, nw)
    flagged = np.nonzero(c["flags"] & 4)[0]
    for out_mode in (1, 2):
        out_hits = np.zeros(n, dtype=HIT_DTYPE) if out_mode == 1 else np.zeros(2 * n, dtype=np.int32)
        hm = np.zeros((n + 31) // 32 + 64, dtype=np.uint32)
        sm = np.zeros_like(hm)
        hs = HostStream(e, 3, 4096, nw)
        e.agg_reset()
        step = chunk // 32 * 32 if out_mode == 2 else chunk
        for k, c0 in enumerate(range(0, n, step)):
            c1 = min(n, c0 + step)
            rows_in = flagged[(flagged >= c0) & (flagged < c1)]
            rn_idx = (rows_in - c0).astype(np.uint32)[::-1].copy()  # (any order)
            rn_rows = np.ascontiguousarray(h_rn[rows_in][::-1])
            slot = k % 3
            hs.wait(slot)
            hs.submit(slot, c1 - c0, h_meta[4 * c0:], h_rows[2 * nw * c0:], nw, c["max_l"], q=h_q[c0:], read_hash=c["rh"][c0:],
                      qname_hash=c["qh"][c0:], idx_base=77 + c0, rn_idx=rn_idx if len(rn_idx) else None,
                      rn_rows=rn_rows if len(rn_idx) else None, emit=True, out_mode=out_mode,
                      out_hits=out_hits[c0:] if out_mode == 1 else out_hits[2 * c0:], out_hit_mask=hm[c0 // 32:], out_strand_mask=sm[c0 // 32:])
        hs.wait_all()
        got_table = e.agg_fetch(e.agg_finalize())
        hs.close()
        assert got_table.tobytes() == want_table.tobytes(), out_mode
        has = (want_hits[:, 2] & 0xFFFF) != 0
        if out_mode == 1:
            assert np.array_equal(out_hits.view(np.uint32).reshape(n, 4), want_hits)
        else:
            bit = lambda m: ((m[np.arange(n) // 32] >> (np.arange(n) % 32).astype(np.uint32)) & 1).astype(bool)  # noqa: E731
            assert np.array_equal(bit(hm), has)
            assert np.array_equal(bit(sm), has & ((want_hits[:, 3] & 1) != 0))
            assert np.array_equal(out_hits.reshape(n, 2)[has], want_hits[:, :2].view(np.int32)[has])
    e.close()


def test_stream_protocol(torch_cuda):
    """a busy slot refuses a second batch; a batch without name hashes needs fragment fields; scan-only batches record nothing"""
    torch = torch_cuda
    from find_circ2_b200._lib import FindCircError, HIT_DTYPE
    from find_circ2_b200.engine import HostStream

    c = _packed_case(torch, 100, 3000, seed=5)
    e, n, nw = c["e"], c["n"], c["nw"]
    h_meta, h_rows, h_q = c["meta"].cpu().numpy(), c["rows"].cpu().numpy(), c["q"].cpu().numpy()
    out = np.zeros(n, dtype=HIT_DTYPE)
    hs = HostStream(e, 2, 4096, nw)
    e.agg_reset()
    with pytest.raises(FindCircError):
        hs.submit(0, 5000, h_meta, h_rows, nw, c["max_l"], emit=False)  # larger than the slots
    hs.submit(0, n, h_meta, h_rows, nw, c["max_l"], emit=False, out_mode=1, out_hits=out)
    with pytest.raises(FindCircError) as ei:
        hs.submit(0, n, h_meta, h_rows, nw, c["max_l"], emit=False)
    assert ei.value.code == -8
    hs.wait(0)
    assert (out["w2"] & 0xFFFF).any() and e.agg_finalize() == 0
    # descriptors packed without fragment fields + no name hashes: the step must fail, not miscount n_frags
    e.agg_reset()
    hs.submit(1, n, h_meta, h_rows, nw, c["max_l"], q=h_q, read_hash=c["rh"], emit=True)
    hs.wait_all()
    with pytest.raises(FindCircError) as ei:
        e.agg_finalize()
    assert ei.value.code == -2
    hs.close()
    e.close()
