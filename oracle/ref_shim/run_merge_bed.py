#!/usr/bin/env python3
"""
Run the reference's own merge_bed.py (Python-2 syntax) inside this python3 container.

TEST INFRASTRUCTURE ONLY -- generates tests/golden/merge/ (tests/golden/make_golden_merge.py).  The reference source is read
from /root/reference/merge_bed.py and never copied; patches applied in memory:
  P1 `print X` -> `print(X)`;  P2 `file(` -> `open(`;  P3 `xrange` -> `range`;  P4 itertools.izip_longest -> zip_longest.
Python-2 semantics that survive unchanged on the inputs used: dict iteration order differs (outputs are compared after a
sort), str() of numpy.float64 / numpy.int64 sums prints the same digits for the integer-valued columns merge_bed.py sums.
"""
import re
import sys

REFERENCE = "/root/reference/merge_bed.py"


def patched_source():
    out = []
    for line in open(REFERENCE).read().split("\n"):
        m = re.match(r"^(\s*)print (.*)$", line)
        if m and not line.lstrip().startswith("#"):
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = re.sub(r"\bfile\(", "open(", line)
        line = re.sub(r"\bxrange\(", "range(", line)
        line = line.replace("from itertools import izip_longest", "from itertools import zip_longest as izip_longest")
        out.append(line)
    return "\n".join(out)


def main():
    sys.argv = [REFERENCE] + sys.argv[1:]
    ns = {"__name__": "__main__", "__file__": REFERENCE}
    exec(compile(patched_source(), REFERENCE, "exec"), ns)
    sys.stdout.flush()
    sfile = ns.get("sfile")  # the statistics file is never closed upstream (python 2 flushed it when the process ended)
    if sfile is not None and sfile is not sys.stderr:
        sfile.close()


if __name__ == "__main__":
    main()
