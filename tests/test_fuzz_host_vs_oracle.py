"""
Differential test on inputs that have no stored reference run: seeded synthetic cases (the generator of tests/golden/:
planted junctions, errors, N's, edge windows, multi-segment and paired fragments) go through the oracle -- which is pinned
to the reference on the goldens -- and through the product's host logic (both ingest paths, fake engine on the CPU).
The GPU twin of this test is tests/test_gpu_pipeline.py on the goldens.
"""
import os
import sys

import pytest

from conftest import GOLDEN
from fake_engine import FakeEngine
from oracle import find_circ_oracle as O

sys.path.insert(0, GOLDEN)
import make_golden as MG  # noqa: E402

CASES = [
    (101, dict(n_pairs=350, read_len=100, asize=15, error_rate=0.01), ["-n", "fz"]),
    (202, dict(n_pairs=300, read_len=76, asize=15, error_rate=0.03), ["-n", "fz", "-m", "3", "-d", "3"]),
    (303, dict(n_pairs=300, read_len=150, asize=20, error_rate=0.02, paired_extra=False), ["-n", "fz", "-a", "20", "--min-uniq-qual", "0"]),
    (404, dict(n_pairs=250, read_len=100, asize=15, error_rate=0.0), ["-n", "fz", "--non-canonical", "--half-unique", "--report-nobridges"]),
    (505, dict(n_pairs=250, read_len=120, asize=18, error_rate=0.015), ["-n", "fz", "-a", "18", "--no-linear"]),
]


@pytest.mark.parametrize("native", [False, True], ids=["python-ingest", "native-ingest"])
@pytest.mark.parametrize("seed,kw,argv", CASES, ids=[str(c[0]) for c in CASES])
def test_product_host_logic_equals_oracle(tmp_path, seed, kw, argv, native):
    from find_circ2_b200 import cli

    case = str(tmp_path)
    MG.build_synth_case(case, seed=seed, **kw)
    fa, sam = os.path.join(case, "genome.fa"), os.path.join(case, "input.sam")
    want = O.run(fa, sam, O.options_from_argv(argv))
    opt = cli.parse_args(["-G", fa] + argv)[0]
    opt.batch_pairs = 173
    eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
    eng.load_genome_fasta(fa)
    out = cli.run_to_strings(opt, sam, engine=eng, native=native)
    assert O.canonical_bed(out["circ"]) == O.canonical_bed(want.circ_bed)
    assert O.canonical_bed(out["lin"]) == O.canonical_bed(want.lin_bed)
    assert out["reads"] == want.reads_fastq
    assert O.canonical_multi(out["multi"]) == O.canonical_multi(want.multi_events)
    assert out["counters"] == want.counters
    assert len(want.circ_bed.splitlines()) > 5
