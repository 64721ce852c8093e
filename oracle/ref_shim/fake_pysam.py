"""
Minimal stand-in for the slice of the pysam API that /root/reference/find_circ.py touches
(find_circ.py:461-469, 473-475, 556-559, 814-817, 842, 898-902, 1027-1056, 1092-1101).

TEST INFRASTRUCTURE ONLY.  It exists so that the UNMODIFIED reference algorithm can be executed
in this container (python2 and pysam are not installable here) by oracle/ref_shim/run_reference.py
in order to generate the golden fixtures under tests/golden/.  Nothing in the product imports it.

Field semantics follow pysam's AlignedRead for SAM text input:
  pos    0-based leftmost coordinate (POS-1)
  aend   0-based exclusive end on the reference = pos + sum(M,D,N,=,X); None when unmapped / no CIGAR
  seq    column 10 as is ('*' -> None)
  qual   column 11 as is ('*' -> None)
  query  seq with leading/trailing soft clips removed (hard clips are not part of seq)
  cigar  list of (op, length), op codes M0 I1 D2 N3 S4 H5 P6 =7 X8 ('*' -> None)
  tags   list of (tag, value); i -> int, f -> float, everything else str
  tid    index of RNAME in @SQ order, -1 for '*'
"""
import re
import sys

_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")
_OPS = {c: i for i, c in enumerate("MIDNSHP=X")}


class AlignedRead(object):
    __slots__ = ("qname", "flag", "tid", "pos", "mapq", "cigar", "seq", "qual", "tags", "_line")

    def __init__(self, line, name2tid):
        f = line.rstrip("\r\n").split("\t")
        self._line = line
        self.qname = f[0]
        self.flag = int(f[1])
        self.tid = name2tid.get(f[2], -1)
        self.pos = int(f[3]) - 1
        self.mapq = int(f[4])
        self.cigar = None if f[5] == "*" else [(_OPS[c], int(n)) for n, c in _CIGAR_RE.findall(f[5])]
        self.seq = None if f[9] == "*" else f[9]
        self.qual = None if f[10] == "*" else f[10]
        tags = []
        for t in f[11:]:
            tag, typ, val = t.split(":", 2)
            if typ == "i":
                val = int(val)
            elif typ == "f":
                val = float(val)
            tags.append((tag, val))
        self.tags = tags

    # FLAG bits
    @property
    def is_paired(self):
        return bool(self.flag & 0x1)

    @property
    def is_unmapped(self):
        return bool(self.flag & 0x4)

    @property
    def is_reverse(self):
        return bool(self.flag & 0x10)

    @property
    def is_read1(self):
        return bool(self.flag & 0x40)

    @property
    def is_read2(self):
        return bool(self.flag & 0x80)

    @property
    def is_secondary(self):
        return bool(self.flag & 0x100)

    @property
    def is_supplementary(self):
        return bool(self.flag & 0x800)

    @property
    def aend(self):
        if self.is_unmapped or not self.cigar:
            return None
        return self.pos + sum(n for op, n in self.cigar if op in (0, 2, 3, 7, 8))

    @property
    def query(self):
        if self.seq is None:
            return None
        s, e = 0, len(self.seq)
        if self.cigar:
            # leading soft clips (hard clips may precede them)
            for op, n in self.cigar:
                if op == 5:
                    continue
                if op == 4:
                    s += n
                else:
                    break
            for op, n in reversed(self.cigar):
                if op == 5:
                    continue
                if op == 4:
                    e -= n
                else:
                    break
        return self.seq[s:e]

    def has_tag(self, tag):
        return any(t == tag for t, v in self.tags)

    def get_tag(self, tag):
        for t, v in self.tags:
            if t == tag:
                return v
        raise KeyError(tag)

    def __str__(self):
        return self._line.rstrip("\n")


class Samfile(object):
    def __init__(self, path, mode="r", template=None):
        if mode not in ("r",):
            raise IOError("fake_pysam only reads SAM text (mode %r requested)" % mode)
        self._fh = sys.stdin if path == "-" else open(path)
        self.references = []
        self.lengths = []
        self._name2tid = {}
        self._pending = None
        # consume header
        for line in self._fh:
            if line.startswith("@"):
                if line.startswith("@SQ"):
                    d = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                    self._name2tid[d["SN"]] = len(self.references)
                    self.references.append(d["SN"])
                    self.lengths.append(int(d["LN"]))
                continue
            self._pending = line
            break

    def getrname(self, tid):
        return self.references[tid]

    def __iter__(self):
        if self._pending is not None:
            line, self._pending = self._pending, None
            yield AlignedRead(line, self._name2tid)
        for line in self._fh:
            if not line.strip():
                continue
            yield AlignedRead(line, self._name2tid)
