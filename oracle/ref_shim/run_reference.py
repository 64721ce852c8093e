#!/usr/bin/env python3
"""
Run the reference's own find_circ.py (v1.99, Python-2 syntax) inside this python3 container.

TEST INFRASTRUCTURE ONLY -- used to generate tests/golden/ (see tests/golden/make_golden.py).
/root/reference does not exist on the GPU box, so nothing at test/bench run time calls this.

How: the reference source is read from /root/reference/find_circ.py (never copied into this repo),
a handful of purely syntactic / python2-semantics patches are applied IN MEMORY, `pysam` is replaced
by oracle/ref_shim/fake_pysam.py (SAM text only), and the module is exec'd with sys.argv set to the
arguments given to this script.  The algorithmic lines (find_circ.py:486-730, 766-974, 976-1140,
1276-1486) run unchanged.

Patches (all python2 -> python3 mechanics, listed so that a reader can audit them):
  P1  `print X`                       -> `print(X)`                         (py2 statement)
  P2  `file(`                         -> `open(`                            (py2 builtin)
  P3  `fromstring(str)`               -> frombuffer(str.encode('latin-1'))  (numpy 2 removed text mode; same bytes)
  P4  `a / b` on ints at :204-205,:590 -> `//`                              (py2 integer division)
  P5  mmap slice is bytes in py3      -> decode latin-1 before .replace()   (:209)
  P6  GzipFile.write(str)             -> wrapper that encodes               (:445, :1447)
  P7  `str(float)`                    -> python2's 12-significant-digit str (:597, :730)
  P8  `from numpy import chararray`   -> dropped (unused, deprecated)
  P9  'string_escape' codec           -> 'unicode_escape'                   (:187, only when an index is re-loaded)
Known behavioural difference that is NOT patched: python2 iterates dicts in hash order, python3 in
insertion order; BED row order therefore differs and all comparisons are made after a canonical sort
(BASELINE.json north_star allows exactly that).
"""
import gzip
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("FIND_CIRC_REFERENCE", "/root/reference/find_circ.py")


def py2_str(x):
    """python2 str(): floats print with 12 significant digits ('%.12g'), always showing a '.0'."""
    if isinstance(x, float):
        s = "%.12g" % x
        if s in ("inf", "-inf", "nan"):
            return s
        if "." not in s and "e" not in s:
            s += ".0"
        return s
    return str(x)


def _make_py2_str_type():
    """A subclass of the builtin named 'str' (optparse inspects type.__name__, find_circ.py:385) whose
    constructor formats floats the python2 way (P7) and returns plain builtin strings."""
    import builtins

    class str(builtins.str):  # noqa: A001 - the name is the point
        def __new__(cls, x=""):
            return builtins.str.__new__(builtins.str, py2_str(x))

    return str


class TextGzip(object):
    """GzipFile that accepts text writes like python2's did (P6). mtime=0 keeps the bytes reproducible."""

    def __init__(self, path, mode="w"):
        self._raw = open(path, "wb")
        self._gz = gzip.GzipFile(filename="", mode="wb", fileobj=self._raw, mtime=0)

    def write(self, s):
        self._gz.write(s.encode("latin-1"))

    def close(self):
        self._gz.close()
        self._raw.close()

    def flush(self):
        self._gz.flush()


def _fromstring(s, dtype=None):
    import numpy as np

    return np.frombuffer(s.encode("latin-1"), dtype=dtype)


def patched_source():
    src = open(REFERENCE).read()
    out = []
    pending_print, depth = False, 0
    for ln, line in enumerate(src.split("\n"), 1):
        m = re.match(r"^(\s*)print (.*)$", line)
        if pending_print:
            # continuation of a print statement whose argument spans lines (find_circ.py:1290-1291)
            depth += line.count("(") - line.count(")")
            if depth == 0:
                line += ")"
                pending_print = False
        elif m and not line.lstrip().startswith("#"):
            depth = m.group(2).count("(") - m.group(2).count(")")
            if depth == 0:
                line = "%sprint(%s)" % (m.group(1), m.group(2))  # P1
            else:
                line = "%sprint(%s" % (m.group(1), m.group(2))
                pending_print = True
        line = re.sub(r"\bfile\(", "open(", line)  # P2
        if line.startswith("from numpy import fromstring,byte"):
            line = "from numpy import byte"  # P3 (fromstring injected below)
        if line.startswith("from numpy import chararray"):
            line = ""  # P8
        if ln in (204, 205):
            assert " / ldata" in line, line
            line = line.replace(" / ldata", " // ldata")  # P4
        if ln == 590:
            assert "len(self.uniq) / 2" in line, line
            line = line.replace("len(self.uniq) / 2", "len(self.uniq) // 2")  # P4
        if ln == 209:
            assert "self.mmap[ofs_start:ofs_end].replace" in line, line
            line = line.replace("self.mmap[ofs_start:ofs_end]", "self.mmap[ofs_start:ofs_end].decode('latin-1')")  # P5
        if ln == 187:
            line = line.replace(".decode('string_escape')", ".encode('latin-1').decode('unicode_escape')")  # P9
        if line.startswith("reads_file = GzipFile("):
            line = line.replace("GzipFile(", "TextGzip(")  # P6
        out.append(line)
    return "\n".join(out)


def main():
    sys.path.insert(0, HERE)
    import fake_pysam

    sys.modules["pysam"] = fake_pysam
    code = compile(patched_source(), REFERENCE, "exec")
    g = {
        "__name__": "__main__",
        "__file__": REFERENCE,
        "str": _make_py2_str_type(),  # P7
        "fromstring": _fromstring,  # P3
        "TextGzip": TextGzip,  # P6
    }
    sys.argv = [REFERENCE] + sys.argv[1:]
    try:
        exec(code, g)
    finally:
        for name in ("circs_file", "lins_file", "reads_file", "multi_file", "test_file"):
            f = g.get(name)
            if f is not None and f is not sys.stdout:
                try:
                    f.close()
                except Exception:
                    pass
        sys.stdout.flush()


if __name__ == "__main__":
    main()
