#!/usr/bin/env python3
"""
Run the reference's own unmapped2anchors.py (Python-2 syntax) inside this python3 container.

TEST INFRASTRUCTURE ONLY -- generates tests/golden/anchors/ (tests/golden/make_golden_anchors.py).  The reference source
is read from /root/reference/unmapped2anchors.py and never copied; patches applied in memory:
  P1  `print X` -> `print(X)`;  P2 `file(` -> `open(`;  P3 numpy.fromstring(str) -> frombuffer of the same bytes;
  pysam -> oracle/ref_shim/fake_pysam.py, with Samfile(path, 'rb') reading SAM TEXT (the fixture is the text form of the
  BAM the reference would be given; the fields used -- qname, flag, seq, qual -- are identical in both encodings).
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference/unmapped2anchors.py"


def patched_source():
    out = []
    for line in open(REFERENCE).read().split("\n"):
        m = re.match(r"^(\s*)print (.*)$", line)
        if m and not line.lstrip().startswith("#"):
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = re.sub(r"\bfile\(", "open(", line)
        line = line.replace("numpy.fromstring(qual,dtype=numpy.uint8)", "numpy.frombuffer(qual.encode('latin-1'),dtype=numpy.uint8)")
        out.append(line)
    return "\n".join(out)


def main():
    sys.path.insert(0, HERE)
    import fake_pysam

    base = fake_pysam.Samfile

    class TextSamfile(base):
        def __init__(self, path, mode="r", template=None):
            base.__init__(self, path, "r")

    fake_pysam.Samfile = TextSamfile
    sys.modules["pysam"] = fake_pysam
    sys.argv = [REFERENCE] + sys.argv[1:]
    exec(compile(patched_source(), REFERENCE, "exec"), {"__name__": "__main__", "__file__": REFERENCE})
    sys.stdout.flush()


if __name__ == "__main__":
    main()
