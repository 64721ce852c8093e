#!/usr/bin/env python3
"""
Fixtures for merge_bed.py: three seeded junction tables in the column layout merge_bed.py's own column map
names (merge_bed.py:117-130: 4 n_reads, 6 n_uniq, 7/8 best_uniq_A/B, 9/10 ov_linear_A/B, 11 samples, 12 per-sample counts,
13-15 edits / anchor_overlap / breakpoints; 16+ free text), merged by the REFERENCE ITSELF through
oracle/ref_shim/run_merge_bed.py in every mode.  Run in the development container (needs /root/reference):
    python tests/golden/make_golden_merge.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "merge")
SHIM = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "ref_shim", "run_merge_bed.py")

HEADER = ["chrom", "start", "end", "name", "n_reads", "strand", "n_uniq", "best_uniq_A", "best_uniq_B", "ov_linear_A", "ov_linear_B",
          "samples", "counts", "edits", "anchor_overlap", "breakpoints", "signal", "strandmatch", "category"]


def table(seed, sample, n, universe):
    rng = np.random.default_rng(seed)
    rows = ["# " + "\t".join(HEADER)]
    pick = rng.choice(len(universe), size=n, replace=False)
    for k, u in enumerate(sorted(pick)):
        chrom, start, end, strand = universe[u]
        n_reads = int(rng.integers(1, 400))
        samples = [sample] if rng.random() < 0.8 else [sample, sample + "_b"]
        counts = [n_reads] if len(samples) == 1 else [n_reads - n_reads // 3, n_reads // 3]
        cats = rng.choice(["CIRCULAR", "CANONICAL", "UNAMBIGUOUS_BP", "PERFECT_EXT", "GOOD_EXT", "ANCHOR_UNIQUE"], size=int(rng.integers(1, 5)), replace=False)
        rows.append("\t".join(str(x) for x in [
            chrom, start, end, "%s_circ_%06d" % (sample, k + 1), n_reads, strand, int(rng.integers(1, n_reads + 1)),
            int(rng.integers(0, 41)), int(rng.integers(0, 41)), int(rng.integers(0, 9)), int(rng.integers(0, 9)),
            ",".join(samples), ",".join(str(c) for c in counts), int(rng.integers(0, 3)), int(rng.integers(0, 3)), int(rng.integers(1, 4)),
            "GTAG" if rng.random() < 0.9 else "GCAG", "NA", ",".join(sorted(cats))]))
    return "\n".join(rows) + "\n"


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(7)
    universe = []
    for _ in range(400):
        c = "chr%d" % rng.integers(1, 6)
        s = int(rng.integers(100, 90000))
        universe.append((c, s, s + int(rng.integers(150, 30000)), "+-"[int(rng.integers(0, 2))]))
    universe = sorted(set(universe))
    names = []
    for k, (sample, n) in enumerate((("liver", 220), ("brain", 180), ("hek", 260), ("heart", 90))):
        p = os.path.join(OUT, "in%d_%s.bed" % (k, sample))
        open(p, "w").write(table(100 + k, sample, n, universe))
        names.append(os.path.basename(p))
    runs = {"default3": names[:3], "default4": names, "bed6": ["-6"] + names[:3], "score": ["--score"] + names, "verbatim": ["-V"] + names[:2],
            "single": names[:1]}
    for tag, argv in runs.items():
        stats = os.path.join(OUT, "ref_%s.stats" % tag)
        res = subprocess.run([sys.executable, SHIM, "-s", os.path.basename(stats)] + argv, cwd=OUT, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        open(os.path.join(OUT, "ref_%s.out" % tag), "w").write(res.stdout)
        open(os.path.join(OUT, "ref_%s.cmd" % tag), "w").write(" ".join(argv) + "\n")
        print(tag, len(res.stdout.splitlines()), "rows")


if __name__ == "__main__":
    main()
