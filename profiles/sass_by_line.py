#!/usr/bin/env python3
"""Join an `ncu --page source --csv` SASS dump with nvdisasm -g line info: executed warp-instructions and stall
samples per source line.  usage: sass_by_line.py <ncu_source.csv> <nvdisasm -g -c output> <kernel substring> [top]"""
import csv
import re
import sys

src_csv, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# line info per instruction offset
off2line = {}
cur = None
infn = False
for line in open(dis):
    if line.startswith(".text.") and line.rstrip().endswith(":"):
        infn = kern in line
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", line)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ie, sm, te = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
base = None
agg = {}
tot = tots = 0
for r in rows[hi + 1 :]:
    if len(r) <= ie or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    ln, ins = off2line.get(a - base, ("?", ""))
    v, s, t = int(r[ie]), int(r[sm]), int(r[te])
    x = agg.setdefault(ln, [0, 0, 0, 0])
    x[0] += v
    x[1] += s
    x[2] += t
    x[3] += 1
    tot += v
    tots += s
print("total warp-instructions %d, samples %d" % (tot, tots))
for ln, (v, s, t, k) in sorted(agg.items(), key=lambda kv: -kv[1][int(__import__("os").environ.get("SORTCOL","0"))])[:top]:
    print("%10d %5.1f%%  samples %5.1f%%  lanes %4.1f  sass %3d  %s" % (v, 100.0 * v / tot, 100.0 * s / max(tots, 1), t / max(v, 1), k, ln))
