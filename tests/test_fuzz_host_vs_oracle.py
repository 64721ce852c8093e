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


def _recombine(sam_in: str, sam_out: str, seed: int):
    """mate pairs made of two fragments of the input each (first fragment = read 1, second = read 2, one name), with the
    order of the records inside a mate shuffled, XS tags dropped at random and some fragments left single: every shape
    the native ingest handles itself (and a few it leaves to python) in one file"""
    import random

    rng = random.Random(seed)
    head, frags = [], []
    for line in open(sam_in):
        if line.startswith("@"):
            head.append(line)
            continue
        f = line.rstrip("\n").split("\t")
        if int(f[1]) & 4:
            continue
        if frags and frags[-1][0][0] == f[0]:
            frags[-1].append(f)
        else:
            frags.append([f])
    rng.shuffle(frags)
    out = list(head)
    k = n = 0
    while k < len(frags):
        name = "fz%d" % n
        n += 1
        mates = [frags[k]] if rng.random() < 0.25 or k + 1 == len(frags) else [frags[k], frags[k + 1]]
        k += len(mates)
        if rng.random() < 0.1 and k < len(frags):  # a third group of records: read 1 again
            mates.append(frags[k])
            k += 1
        for m, recs in enumerate(mates):
            recs = list(recs)
            if rng.random() < 0.3:
                rng.shuffle(recs)
            for f in recs:
                f = list(f)
                f[0] = name
                f[1] = str((int(f[1]) & ~0xC1) | ((0x41 if m % 2 == 0 else 0x81) if len(mates) > 1 else 0))
                if rng.random() < 0.2:
                    f = [x for x in f if not x.startswith("XS:i:")]
                out.append("\t".join(f) + "\n")
        if rng.random() < 0.05:
            out.append("um%d\t4\t*\t0\t0\t*\t*\t0\t0\tACGTACGTAC\t*\n" % n)
    open(sam_out, "w").write("".join(out))


@pytest.mark.parametrize("seed,argv", [(1, ["-n", "fz"]), (2, ["-n", "fz", "--no-linear"]), (3, ["-n", "fz", "--min-uniq-qual", "8"]),
                                       (4, ["-n", "fz", "--no-multi", "--half-unique", "--report-nobridges"])])
def test_native_evidence_rules_equal_python_on_recombined_mates(tmp_path, seed, argv):
    """the batch-wide evidence rules of the native path (pipeline._native_batch) against the per-fragment reading of
    record_hits (pipeline._record_hits, pinned to the reference by the goldens), and both against the oracle"""
    from find_circ2_b200 import cli

    case = str(tmp_path)
    MG.build_synth_case(case, seed=40 + seed, n_pairs=500, read_len=100, asize=15, error_rate=0.01)
    fa, sam = os.path.join(case, "genome.fa"), os.path.join(case, "mates.sam")
    _recombine(os.path.join(case, "input.sam"), sam, seed)
    want = O.run(fa, sam, O.options_from_argv(argv))
    outs = []
    for native in (False, True):
        opt = cli.parse_args(["-G", fa] + argv)[0]
        opt.batch_pairs = 97
        eng = FakeEngine(0, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
        eng.load_genome_fasta(fa)
        outs.append(cli.run_to_strings(opt, sam, engine=eng, native=native))
    py, nat = outs
    for k in ("circ", "lin", "reads", "multi", "counters"):
        assert nat[k] == py[k], k
    assert O.canonical_bed(nat["circ"]) == O.canonical_bed(want.circ_bed)
    assert O.canonical_bed(nat["lin"]) == O.canonical_bed(want.lin_bed)
    assert nat["reads"] == want.reads_fastq
    assert O.canonical_multi(nat["multi"]) == O.canonical_multi(want.multi_events)
    assert nat["counters"] == want.counters
    flags = set(",".join(ln.split("\t")[20] for ln in nat["circ"].splitlines()[1:]).split(","))
    assert {"WARN_MULTI_BACKSPLICE", "SUPPORT_CLOSURE", "WARN_OUTSIDE_MATE", "WARN_OTHER_CHROM_MATE"} <= flags or seed != 1
