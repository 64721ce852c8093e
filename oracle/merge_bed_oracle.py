"""
CPU restatement of the reference's merge_bed.py (TEST INFRASTRUCTURE: only tests/ may import it).

Follows /root/reference/merge_bed.py line by line in python 3:
  read_to_hash        :25-58    one dict per input, key (chrom, start, end, sense); a later row with the same key replaces
                                an earlier one of the same file
  support / stats     :62-91    which inputs hold a key; "<in_i>_AND_<in_j>\\t<count>" lines, sorted
  consensus_cols      :93-142   column map: 3 names joined sorted, 4 / 6 float sums, 7 / 8 max, 9 / 10 int sums, 11 sample lists
                                concatenated and sorted, 12 per-sample counts summed (samples sorted), 13-15 min, every other
                                column the sorted union of its comma-separated tokens (empty tokens dropped)
  output modes        :144-181  verbatim, --score, default
Pinned by tests/golden/merge/ref_*.out, produced by the reference itself (oracle/ref_shim/run_merge_bed.py).
Row order is python-2 dict order upstream: compare after a sort.  In --score mode the name column is ",".join(set(...)):
compare it as a set.
"""
from collections import Counter, defaultdict
from itertools import zip_longest


def py2_float_str(x: float) -> str:
    s = "%.12g" % x
    return s + ".0" if ("." not in s and "e" not in s and s not in ("inf", "-inf", "nan")) else s


def read_table(path, bed6=False):
    pos = {}
    for line in open(path):
        if line.startswith("#"):
            continue
        parts = line.strip().split("\t")
        if bed6:
            parts = parts[:6]
        chrom, start, end, name, score, sense = parts[:6]
        pos[(chrom, int(start), int(end), sense)] = parts
    return pos


def consensus_cols(lines):
    samples = []
    counts = defaultdict(int)

    def setup_samples(values):
        alls = []
        for v in values:
            toadd = v.split(",")
            samples.append(toadd)
            alls.extend(toadd)
        return ",".join(sorted(alls))

    def assign_counts(values):
        for cs, ss in zip(values, samples):
            for samp, count in zip(ss, cs.split(",")):
                counts[samp] += int(count)
        return ",".join(str(counts[k]) for k in sorted(counts))

    def append_uniq(values):
        v = set()
        for row in values:
            v |= set(row.split(","))
        return ",".join(str(x) for x in sorted(v) if x)

    col_map = {
        3: lambda v: ",".join(sorted(v)),
        4: lambda v: py2_float_str(sum(float(x) for x in v)),
        6: lambda v: py2_float_str(sum(float(x) for x in v)),
        7: lambda v: max(int(x) for x in v),
        8: lambda v: max(int(x) for x in v),
        9: lambda v: sum(int(x) for x in v),
        10: lambda v: sum(int(x) for x in v),
        11: setup_samples,
        12: assign_counts,
        13: lambda v: min(int(x) for x in v),
        14: lambda v: min(int(x) for x in v),
        15: lambda v: min(int(x) for x in v),
    }
    parts = []
    for i, column in enumerate(zip_longest(*lines, fillvalue="")):
        parts.append(str(col_map[i](column)) if i in col_map else append_uniq(column))
    return parts


def merge(paths, bed6=False, score=False, verbatim=False):
    """returns (output text, stats text)"""
    inputs = [read_table(p, bed6) for p in paths]
    shorts = ["in%d" % i for i in range(len(paths))]
    by_name = dict(zip(shorts, inputs))
    merged = {}
    support = defaultdict(list)
    for name, data in zip(shorts, inputs):
        merged.update(data)
        for pos in data:
            support[pos].append(name)
    comb = Counter(tuple(v) for v in support.values())
    stats = "".join("%s\t%d\n" % ("_AND_".join(c), comb[c]) for c in sorted(comb))
    out = []
    for pos in merged:
        com = support[pos]
        comstr = "(%s)" % ",".join(com)
        if verbatim:
            cols = [comstr]
            for name in com:
                cols.append("%s : " % name)
                cols.append("\t".join(by_name[name][pos]))
        elif score:
            cols = [",".join(set(by_name[name][pos][3] for name in com))]
            for name in shorts:
                cols.append(by_name[name][pos][4] if name in com else "0")
            cols.append(comstr)
        else:
            cols = [comstr] + consensus_cols([by_name[name][pos] for name in com])
        out.append("\t".join(cols) + "\n")
    return "".join(out), stats


def canonical(text, score=False):
    rows = []
    for line in text.split("\n"):
        if not line:
            continue
        c = line.split("\t")
        if score:
            c[0] = ",".join(sorted(c[0].split(",")))
        rows.append("\t".join(c))
    return sorted(rows)
