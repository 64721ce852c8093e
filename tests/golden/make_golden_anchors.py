#!/usr/bin/env python3
"""
Golden fixtures for the anchor cutter (find_circ2_b200/unmapped2anchors.py): a SAM text file of mostly unmapped reads
with qualities around the filter's threshold (including characters below '#', which wrap in the reference's uint8
arithmetic) and the output of the REFERENCE's unmapped2anchors.py on it (oracle/ref_shim/run_unmapped2anchors.py) for
the deterministic option sets.  Run in the dev container only:  python tests/golden/make_golden_anchors.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "anchors")
SHIM = os.path.join(ROOT, "oracle", "ref_shim", "run_unmapped2anchors.py")
OPTION_SETS = {"default": [], "a15_q20": ["-a", "15", "-q", "20"], "revA": ["-r", "A"], "revB": ["-r", "B"], "revR": ["-r", "R"],
               "revC": ["-r", "C"], "a25_q0": ["-a", "25", "-q", "0"]}


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(5)
    lines = ["@HD\tVN:1.0\tSO:unsorted\n", "@SQ\tSN:chr1\tLN:100000\n"]
    for k in range(300):
        L = int(rng.choice([50, 76, 100, 101, 150]))
        seq = "".join(rng.choice(list("ACGT"), size=L))
        if k % 17 == 0:
            seq = seq[:7] + "N" + seq[8:]
        if k % 23 == 0:
            seq = seq.lower()
        kind = k % 5
        if kind == 0:
            q = rng.integers(35, 74, L)                      # good
        elif kind == 1:
            q = np.concatenate([rng.integers(35, 42, 25), rng.integers(60, 74, L - 25)])  # weak head
        elif kind == 2:
            q = np.concatenate([rng.integers(60, 74, L - 25), rng.integers(35, 42, 25)])  # weak tail
        elif kind == 3:
            q = rng.integers(33, 45, L)                      # '!'..: below '#' wraps around in uint8
        else:
            q = rng.integers(38, 43, L)                      # right at the default threshold
        qual = "".join(chr(int(c)) for c in q)
        flag = 4 if k % 7 else 0                             # every 7th record is mapped and must be skipped
        rname, pos, cigar = ("*", 0, "*") if flag else ("chr1", 1000 + k, "%dM" % L)
        lines.append("read%d/%d\t%d\t%s\t%d\t0\t%s\t*\t0\t0\t%s\t%s\n" % (k, k % 2 + 1, flag, rname, pos, cigar, seq, qual))
    with open(os.path.join(OUT, "unmapped.sam"), "w") as fh:
        fh.writelines(lines)
    for tag, args in OPTION_SETS.items():
        r = subprocess.run([sys.executable, SHIM] + args + [os.path.join(OUT, "unmapped.sam")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        with open(os.path.join(OUT, "ref_%s.fastq" % tag), "w") as fh:
            fh.write(r.stdout)
        with open(os.path.join(OUT, "ref_%s.args" % tag), "w") as fh:
            fh.write(" ".join(args) + "\n")
        print(tag, r.stdout.count("\n") // 8, "reads kept")


if __name__ == "__main__":
    main()
