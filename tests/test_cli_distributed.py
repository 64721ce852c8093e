"""
The drop-in on N processes (find_circ2_b200.cli.run_distributed): every rank parses its own byte range of the SAM file
(cut where the read name changes), junction records are exchanged by key, rank 0 gathers junction rows and the host-side
companions and writes -- the outputs must be the single-process outputs, i.e. the reference's (find_circ.py:1605-1610).
CPU: gloo, world sizes 2 and 3, the GPU engine replaced by the test double; the GPU twin runs under torchrun in
tests/multigpu_check.py and in bench.py.
"""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, golden_cases

sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [(c, r, a) for c, r, a in golden_cases()
         if not any(x in a for x in ("--all-hits", "--noop", "--test", "--stdout")) and os.path.basename(r) in
         ("ref_default", "ref_a20", "ref_known", "ref_uniq0_half_nobridge", "ref_nolinear_nomulti", "ref_m4_d3")]


def _worker(rank, world, port, case_dir, ref_dir, argv, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import helpers as H
    from fake_engine import FakeEngine
    from find_circ2_b200 import cli

    dist, dev = cli.distributed_context(device_ok=False)
    opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv)[0]
    opt.batch_pairs = 97
    eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
    eng.load_genome_fasta(opt.genome)
    out = cli.run_distributed(opt, os.path.join(case_dir, "input.sam"), dist, dev, engine=eng)
    if rank == 0:
        H.compare_outputs(out["circ"], out["lin"], out["reads"], out["multi"], out["counters"], ref_dir, argv)
        open(out_path, "w").write("ok %d" % out["n_fragments"])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


RUNS = [(2, c, r, a) for c, r, a in CASES] + [(3, c, r, a) for c, r, a in CASES if os.path.basename(r) == "ref_default"]


@pytest.mark.parametrize("world,case_dir,ref_dir,argv", RUNS,
                         ids=["w%d-%s-%s" % (w, os.path.basename(c), os.path.basename(r)[4:]) for w, c, r, _ in RUNS])
def test_distributed_dropin_reproduces_reference_runs(tmp_path, world, case_dir, ref_dir, argv):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "done.txt")
    mp.spawn(_worker, args=(world, port, case_dir, ref_dir, argv, out_path), nprocs=world, join=True)
    assert open(out_path).read().startswith("ok")


def test_ranges_cut_on_fragment_boundaries(tmp_path):
    from find_circ2_b200 import cli

    sam = os.path.join(GOLDEN, "synth_a", "input.sam")
    body_lines = [ln for ln in open(sam, "rb").read().split(b"\n") if ln and not ln.startswith(b"@")]
    for world in (1, 2, 5, 64):
        names, ranges, body = cli.sam_ranges(sam, world)
        assert ranges[0][0] == body and ranges[-1][1] == os.path.getsize(sam)
        data = open(sam, "rb").read()
        seen = []
        for (s, e), nxt in zip(ranges, ranges[1:] + [None]):
            assert s <= e and (nxt is None or nxt[0] == e)
            part = [ln for ln in data[s:e].split(b"\n") if ln]
            seen += part
            if part and e < len(data):
                assert data[e:].split(b"\t", 1)[0] != part[-1].split(b"\t", 1)[0]  # no read name straddles a cut
        assert seen == body_lines
