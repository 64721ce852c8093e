"""
merge_bed.py of the reference (merge several find_circ junction tables on their genomic coordinates,
/root/reference/merge_bed.py) with the same command line and output -- the keyed merge runs on the GPU.

    merge_bed.py [-s stats] [-6] [--score] [-V] 1.bed 2.bed [3.bed ...] > merged.bed

Host: reads the tables (a later row of one file replaces an earlier one with the same key, merge_bed.py:45), hands key and
numeric columns of ALL rows to the device (fc_merge_tables: group by (chrom, start, end, strand), per group the inputs that
hold it, sums / maxima / minima of the column map merge_bed.py:117-130), then joins the text columns of each group's rows and
prints.  Output rows come in key order (upstream: python-2 dict order; compare after a sort); `-f/--flank` is accepted and
ignored exactly as upstream ignores it (merge_bed.py:64 always passes flank=0).  At most 64 input files per call.
"""
from __future__ import annotations

import ctypes as C
import optparse
import os
import sys
from collections import Counter, defaultdict

import numpy as np

USAGE = """
%prog 1.bed 2.bed [3.bed] [4.bed] [...] > merged.bed

Merge BED or BED-like files on the genomic coordinates. Deals properly
with find_circ.py output and adds a few extra columns.
"""

# merge_bed.py:117-130 -- column -> (reduction on the device, how the result prints)
SUM, MAX, MIN = 0, 1, 2
NUMERIC = {4: (SUM, "float"), 6: (SUM, "float"), 7: (MAX, "int"), 8: (MAX, "int"), 9: (SUM, "int"), 10: (SUM, "int"),
           13: (MIN, "int"), 14: (MIN, "int"), 15: (MIN, "int")}


def py2_float_str(x: float) -> str:
    s = "%.12g" % x
    return s + ".0" if ("." not in s and "e" not in s and s not in ("inf", "-inf", "nan")) else s


def read_table(path: str, bed6: bool):
    """rows of one input, one per key, in first-appearance order of the keys (merge_bed.py:25-58)"""
    pos = {}
    fh = sys.stdin if path == "-" else open(path)
    try:
        for line in fh:
            if line.startswith("#"):
                continue
            parts = line.strip().split("\t")
            if bed6:
                parts = parts[:6]
            chrom, start, end, name, score, sense = parts[:6]
            pos[(chrom, int(start), int(end), sense)] = parts
    finally:
        if fh is not sys.stdin:
            fh.close()
    return pos


def device_merge(inputs, n_cols_of_interest, engine=None):
    """group the rows of all inputs on the GPU; returns (groups: list of row-index lists in input order, support masks,
    reduced numeric columns {col: array}, flat row list)"""
    from . import _lib

    lib = _lib.load()
    rows, src = [], []
    for k, data in enumerate(inputs):
        for key, parts in data.items():
            rows.append((key, parts))
            src.append(k)
    n = len(rows)
    chrom_ids = {}
    chrom = np.zeros(n, dtype=np.uint32)
    start = np.zeros(n, dtype=np.int32)
    end = np.zeros(n, dtype=np.int32)
    strand = np.zeros(n, dtype=np.uint8)
    strands = {}
    for i, ((c, s, e, sense), _) in enumerate(rows):
        chrom[i] = chrom_ids.setdefault(c, len(chrom_ids))
        start[i], end[i] = s, e
        strand[i] = strands.setdefault(sense, len(strands))
    if len(strands) > 255:
        raise ValueError("more than 255 distinct values in the strand column")
    cols = sorted(c for c in n_cols_of_interest)
    vals = np.zeros((max(len(cols), 1), n), dtype=np.float64)
    for j, c in enumerate(cols):
        for i, (_, parts) in enumerate(rows):
            vals[j, i] = float(parts[c]) if NUMERIC[c][1] == "float" else int(parts[c])
    ops = np.array([NUMERIC[c][0] for c in cols] or [0], dtype=np.uint8)
    own = engine is None
    h = C.c_void_p()
    if own:
        rc = lib.fc_ctx_create(0, C.byref(h))
        if rc != 0:
            raise RuntimeError(lib.fc_last_error(None).decode())
    else:
        h = engine.h
    try:
        ng = C.c_int64(0)
        group_of_row = np.zeros(max(n, 1), dtype=np.uint32)
        support = np.zeros(max(n, 1), dtype=np.uint64)
        out = np.zeros((max(len(cols), 1), max(n, 1)), dtype=np.float64)
        flat_out = np.zeros(max(len(cols), 1) * max(n, 1), dtype=np.float64)
        rc = lib.fc_merge_tables(h, n, chrom.ctypes.data, start.ctypes.data, end.ctypes.data, strand.ctypes.data,
                                 np.asarray(src, dtype=np.uint8).ctypes.data, len(cols), vals.ctypes.data, ops.ctypes.data, C.byref(ng),
                                 group_of_row.ctypes.data, support.ctypes.data, flat_out.ctypes.data)
        if rc != 0:
            raise RuntimeError(lib.fc_last_error(h).decode())
    finally:
        if own:
            lib.fc_ctx_destroy(h)
    ng = int(ng.value)
    out = flat_out[: len(cols) * ng].reshape(len(cols), ng) if cols else np.zeros((0, ng))
    groups = [[] for _ in range(ng)]
    for i in range(n):
        groups[int(group_of_row[i])].append(i)  # ascending row index = input order
    return groups, support[:ng], {c: out[j] for j, c in enumerate(cols)}, rows, src


def merge_to_text(paths, bed6=False, score=False, verbatim=False, engine=None):
    """(output text, stats text) of merge_bed.py for these inputs"""
    if len(paths) > 64:
        raise SystemExit("at most 64 input files per call")
    inputs = [read_table(p, bed6) for p in paths]
    shorts = ["in%d" % i for i in range(len(paths))]
    numeric = {} if (verbatim or score) else {c: v for c, v in NUMERIC.items() if not bed6 or c < 6}
    # a column of the map only exists when every row is that wide; narrower tables keep it as text, as zip_longest would
    width = min((len(p) for d in inputs for p in d.values()), default=0)
    numeric = {c: v for c, v in numeric.items() if c < width}
    groups, support, red, rows, src = device_merge(inputs, numeric, engine)
    comb = Counter()
    out = []
    for g, members in enumerate(groups):
        com = [shorts[k] for k in range(len(paths)) if (int(support[g]) >> k) & 1]
        comb[tuple(com)] += 1
        comstr = "(%s)" % ",".join(com)
        lines = [rows[i][1] for i in members]
        if verbatim:
            cols = [comstr]
            for i in members:
                cols.append("%s : " % shorts[src[i]])
                cols.append("\t".join(rows[i][1]))
        elif score:
            have = {shorts[src[i]]: rows[i][1][4] for i in members}
            cols = [",".join(dict.fromkeys(ln[3] for ln in lines))] + [have.get(s, "0") for s in shorts] + [comstr]
        else:
            cols = [comstr] + consensus_text(lines, {c: red[c][g] for c in numeric}, numeric)
        out.append("\t".join(cols) + "\n")
    stats = "".join("%s\t%d\n" % ("_AND_".join(c), comb[c]) for c in sorted(comb))
    return "".join(out), stats


def consensus_text(lines, reduced, numeric):
    """consensus_cols (merge_bed.py:93-142): numeric columns come reduced from the device, text columns are joined here"""
    from itertools import zip_longest

    samples = []
    counts = defaultdict(int)
    parts = []
    for i, column in enumerate(zip_longest(*lines, fillvalue="")):
        if i in numeric:
            v = float(reduced[i])
            parts.append(py2_float_str(v) if numeric[i][1] == "float" else str(int(v)))
        elif i == 3:
            parts.append(",".join(sorted(column)))
        elif i == 11:
            alls = []
            for v in column:
                toadd = v.split(",")
                samples.append(toadd)
                alls.extend(toadd)
            parts.append(",".join(sorted(alls)))
        elif i == 12:
            for cs, ss in zip(column, samples):
                for samp, count in zip(ss, cs.split(",")):
                    counts[samp] += int(count)
            parts.append(",".join(str(counts[k]) for k in sorted(counts)))
        else:
            v = set()
            for row in column:
                v |= set(row.split(","))
            parts.append(",".join(str(x) for x in sorted(v) if x))
    return parts


def main(argv=None) -> int:
    p = optparse.OptionParser(usage=USAGE)
    p.add_option("-f", "--flank", dest="flank", type=int, default=0, help="accepted, without effect (as upstream)")
    p.add_option("-s", "--stats", dest="stats", default="", help="write statistics to this file (instead of stderr)")
    p.add_option("-6", "--bed6", dest="bed6", default=False, action="store_true", help="ignore all columns except the first six")
    p.add_option("", "--score", dest="score", default=False, action="store_true", help="name, the score of every input, support")
    p.add_option("-F", "--format", dest="format", default="2", choices=["1", "1.2", "2"], help="accepted, without effect (as upstream)")
    p.add_option("-V", "--verbatim", dest="verbatim", default=False, action="store_true", help="join on coordinates, other columns verbatim")
    o, args = p.parse_args(sys.argv[1:] if argv is None else list(argv))
    text, stats = merge_to_text(args, bed6=o.bed6, score=o.score, verbatim=o.verbatim)
    if o.stats:
        with open(o.stats, "w") as fh:
            fh.write(stats)
    else:
        sys.stderr.write(stats)
    sys.stdout.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
