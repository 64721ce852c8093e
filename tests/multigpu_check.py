#!/usr/bin/env python3
"""
Multi-GPU parity check, run by hand on a GPU box (it needs torchrun and >= 2 GPUs, so pytest does not collect it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

Every rank scans its own shard of a seeded workload; the junction table is built three ways and must be identical:
  (a) fused emit + exchange over peer memory (CUDA IPC)          find_circ2_b200/parallel.py: p2p_setup / fc_agg_emit_p2p
  (b) partition + NCCL all-to-all                                find_circ2_b200/parallel.py: exchange_records
  (c) rank 0 alone aggregating the union of all ranks' records   (single-GPU path)
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import helpers as H
    from find_circ2_b200 import parallel, synth
    from find_circ2_b200._lib import JREC_DTYPE
    from find_circ2_b200.engine import Engine

    g = synth.make_genome([400000, 300000], seed=1, n_frac=0.005, n_run=(20, 300))
    J = synth.plant_junctions(g, 300, 60, seed=2, span=(150, 20000), margin=300)
    t = synth.make_pairs(g, J, 60000, read_len=100, asize=20, seed=10 + rank, error_rate=0.01)
    chrom, a_start, b_end, l, flags, internal = H.pairs_to_soa(t, 20, 2)
    n = len(chrom)
    e = Engine(device=local, asize=20)
    e.load_genome_arrays(g.names, g.seqs)
    rh = e.hash_reads(t.reads, np.full(n, 100, dtype=np.int32))
    qh = (t.name_id.astype(np.uint64) + np.uint64(rank << 32)) * np.uint64(0x9E3779B97F4A7C15)
    qa = (t.as_a - np.maximum(t.xs_a, 0)).astype(np.int16)
    qb = (t.as_b - np.maximum(t.xs_b, 0)).astype(np.int16)
    tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_chrom, d_a, d_b, d_l, d_fl, d_asc = tn(chrom), tn(a_start), tn(b_end), tn(l), tn(flags), tn(internal)
    n_words = Engine.n_words_for(int(l.max()))
    planes = torch.zeros(3 * n_words * n, dtype=torch.int32, device=dev)
    hits = torch.zeros(n * 4, dtype=torch.int32, device=dev)
    e.pack_reads(d_asc, internal.shape[1], d_l, n_words, planes, d_fl, 0)
    pairs = e.make_pairs(n, d_chrom, d_a, d_b, d_l, d_fl, planes, n_words, int(l.max()))
    e.scan(pairs, hits, 0)
    pay = (tn(np.ones(n, np.uint8)), tn(qa), tn(qb), tn(rh.view(np.int64)), tn(qh.view(np.int64)))
    # stream positions: every rank holds one contiguous shard of the input stream
    n_cap = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_reduce(n_cap, op=dist.ReduceOp.MAX)
    n_cap = int(n_cap.item())
    idx_base = rank * n_cap

    def table():
        nj = e.agg_finalize(0)
        return parallel.gather_junctions(e.agg_fetch(nj), dist, dev)

    # (b) NCCL all-to-all
    e.agg_reset()
    e.agg_emit(n, hits, d_chrom, d_fl, *pay, idx_base, 0)
    parallel.exchange_records(e, dist, dev, 0, upper_bound=n)
    tb = table()

    # (a) peer-memory emit
    ok = parallel.p2p_setup(e, dist, dev, int(2.5 * n) + 4096)
    if not ok:
        if rank == 0:
            print("P2P (CUDA IPC) not available on this box -- only the NCCL path was checked")
        ta = tb
    else:
        for it in range(4):  # repeated steps: the two halves of the peer buffers alternate, one barrier per step
            e.agg_reset_async(0)
            if it > 0:  # declared idx range: discovery rank from flags; undeclared (it == 0): from a sort
                e.agg_set_idx_range(0, world * n_cap)
            if it % 2 == 1:  # scan and peer emit in one kernel
                hits2 = torch.zeros_like(hits)
                e.scan_emit_p2p(pairs, hits2, *pay, idx_base, 0)
                assert torch.equal(hits, hits2)
            else:
                e.agg_emit_p2p(n, hits, d_chrom, d_fl, *pay, idx_base, 0)
            parallel.stream_barrier(dist, dev, e, 0)  # ends the step: slice counts published, stores ordered
            ta = table()

    # (c) union on rank 0: gather every rank's hits + payload and aggregate alone
    h_hits = hits.cpu().numpy()
    parts = [None] * world
    dist.all_gather_object(parts, dict(hits=h_hits, chrom=chrom, flags=flags, qa=qa, qb=qb, rh=rh, qh=qh, base=idx_base))
    if rank == 0:
        # (a context that is connected to peers reduces what its peers send: the union goes through a second context)
        e2 = Engine(device=local, asize=20)
        e2.share_genome(e)
        e2.agg_reset()
        for p in parts:
            m = len(p["chrom"])
            e2.agg_emit(m, tn(p["hits"]), tn(p["chrom"]), tn(p["flags"]), tn(np.ones(m, np.uint8)), tn(p["qa"]), tn(p["qb"]),
                        tn(p["rh"].view(np.int64)), tn(p["qh"].view(np.int64)), p["base"], 0)
        nj = e2.agg_finalize(0)
        tc = e2.agg_fetch(nj)
        e2.close()
        assert len(tc) > 100
        assert ta.tobytes() == tc.tobytes(), "peer-memory path differs from the single-GPU aggregation"
        assert tb.tobytes() == tc.tobytes(), "NCCL path differs from the single-GPU aggregation"
        print("multigpu_check OK: world=%d, %d junctions, p2p=%s" % (world, len(tc), ok))
    dist.barrier()
    # the whole drop-in on `world` GPUs against runs of the reference itself (tests/golden)
    from conftest import golden_cases
    from find_circ2_b200 import cli

    done = 0
    for case_dir, ref_dir, argv in golden_cases():
        if os.path.basename(ref_dir) not in ("ref_default", "ref_a20", "ref_uniq0_half_nobridge") or "--stdout" in argv:
            continue
        opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv)[0]
        opt.device, opt.batch_pairs = local, 257
        out = cli.run_distributed(opt, os.path.join(case_dir, "input.sam"), dist, dev)
        if rank == 0:
            H.compare_outputs(out["circ"], out["lin"], out["reads"], out["multi"], out["counters"], ref_dir, argv)
            done += 1
        dist.barrier()
    if rank == 0:
        print("multigpu_check OK: drop-in on %d GPUs reproduces %d reference runs" % (world, done))
    dist.destroy_process_group()
    e.close()


if __name__ == "__main__":
    main()
