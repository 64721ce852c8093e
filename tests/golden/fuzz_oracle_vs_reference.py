#!/usr/bin/env python3
"""
Wider net than the stored goldens: random seeds x random pairs of options, each run through the REFERENCE
(oracle/ref_shim/run_reference.py) and through the oracle, all five outputs compared.  Dev container only (needs
/root/reference); nothing is stored.      python tests/golden/fuzz_oracle_vs_reference.py [seed] [runs]

Round 1: 110 runs (seeds 3 and 99), no difference.
"""
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as MG  # noqa: E402
from oracle import find_circ_oracle as O  # noqa: E402

OPTIONS = [[], ["-m", "3"], ["-d", "1"], ["-d", "3", "-m", "4"], ["--non-canonical"], ["--half-unique"], ["--report-nobridges"],
           ["--min-uniq-qual", "0"], ["--no-linear"], ["--no-multi"], ["--strand-pref"], ["--all-hits"],
           ["--short-threshold", "300", "--huge-threshold", "2000"], ["-d", "0"], ["-m", "0"]]


def main():
    rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    done = bad = 0
    for _ in range(runs):
        seed = rnd.randrange(1000, 100000)
        read_len, asize = rnd.choice([60, 76, 100, 125, 150]), rnd.choice([12, 15, 18, 20])
        if read_len < 3 * asize:
            continue
        o1, o2 = rnd.choice(OPTIONS), rnd.choice(OPTIONS)
        if o1 and o2 and o1[0] == o2[0]:
            o2 = []
        argv = ["-n", "fz", "-a", str(asize)] + o1 + o2
        with tempfile.TemporaryDirectory() as case:
            try:
                MG.build_synth_case(case, seed=seed, n_pairs=rnd.choice([150, 250]), read_len=read_len, asize=asize,
                                    error_rate=rnd.choice([0, 0.01, 0.03]), paired_extra=rnd.random() < 0.7)
            except Exception:
                continue  # the generator could not place its fragments for this geometry
            MG.run_reference(case, "x", argv)
            ref = os.path.join(case, "ref_x")
            if "exit=0" not in open(os.path.join(ref, "cmdline.txt")).read():
                print("reference failed", argv)
                continue
            out = O.run(os.path.join(case, "genome.fa"), os.path.join(case, "input.sam"), O.options_from_argv(argv))
            rd = lambda n: open(os.path.join(ref, n)).read()  # noqa: E731
            same = (O.canonical_bed(out.circ_bed) == O.canonical_bed(rd("circ_splice_sites.bed"))
                    and O.canonical_bed(out.lin_bed) == O.canonical_bed(rd("lin_splice_sites.bed"))
                    and out.reads_fastq == rd("spliced_reads.fastq")
                    and O.canonical_multi(out.multi_events) == O.canonical_multi(rd("multi_events.tsv"))
                    and out.counters == rd("counters.txt"))
            done += 1
            if not same:
                bad += 1
                print("ORACLE != REFERENCE: seed %d read_len %d %s" % (seed, read_len, " ".join(argv)))
    print("%d runs, %d differences" % (done, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
