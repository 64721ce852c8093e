"""
Alignment input: SAM text and BAM (BGZF) decoded on the host into light records with the field semantics the
reference gets from pysam (find_circ.py:461-469 opens the stream, :1450-1486 consumes it):

  pos    0-based leftmost reference coordinate
  aend   0-based exclusive reference end = pos + sum of M/D/N/=/X lengths (None when unmapped or without CIGAR)
  seq    sequence as stored (genome orientation), None for '*'
  qual   quality string as stored, None for '*'
  cigar  list of (op, length) with op codes M0 I1 D2 N3 S4 H5 P6 =7 X8
  AS/XS  integer tags (None when absent)
  tid    index of the reference name in @SQ order (-1 for '*')

pysam is not used (it is not installed on the target image); the BAM reader relies on BGZF being a series of
gzip members, which python's gzip module reads natively.
"""
from __future__ import annotations

import gzip
import re
import struct
import sys
from typing import Iterator, List, Optional, Tuple

_CIG = re.compile(r"(\d+)([MIDNSHP=X])")
_OP = {c: i for i, c in enumerate("MIDNSHP=X")}
_REF_CONSUMING = (0, 2, 3, 7, 8)


class Alignment(object):
    __slots__ = ("qname", "flag", "tid", "pos", "cigar", "seq", "qual", "AS", "XS", "aend", "line_no")

    def __init__(self, qname, flag, tid, pos, cigar, seq, qual, AS, XS, line_no=0):
        self.qname = qname
        self.flag = flag
        self.tid = tid
        self.pos = pos
        self.cigar = cigar
        self.seq = seq
        self.qual = qual
        self.AS = AS
        self.XS = XS
        self.line_no = line_no
        if (flag & 0x4) or not cigar:
            self.aend = None
        else:
            e = pos
            for op, n in cigar:
                if op in _REF_CONSUMING:
                    e += n
            self.aend = e

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

    def clip_start(self) -> int:
        """offset of the aligned part inside the full read: soft and hard clip lengths are summed until the first
        match operation; other operations do not stop the walk (find_circ.py:1086-1097)"""
        s = 0
        for op, n in self.cigar:
            if op == 4 or op == 5:
                s += n
            elif op == 0:
                break
        return s

    def query_length(self) -> int:
        """len(pysam's .query): stored sequence minus soft clips (hard-clipped bases are not stored)"""
        n = len(self.seq)
        cig = self.cigar
        if cig:
            for op, c in cig:
                if op == 5:
                    continue
                if op == 4:
                    n -= c
                else:
                    break
            for op, c in reversed(cig):
                if op == 5:
                    continue
                if op == 4:
                    n -= c
                else:
                    break
        return n

    def uniqueness(self) -> int:
        """AS - XS, AS alone when XS is absent (find_circ.py:809-819)"""
        return self.AS if self.XS is None else self.AS - self.XS


def _parse_sam_line(line: str, name2tid, line_no: int) -> Alignment:
    f = line.rstrip("\r\n").split("\t")
    cigar = None if f[5] == "*" else [(_OP[c], int(n)) for n, c in _CIG.findall(f[5])]
    AS = XS = None
    for t in f[11:]:
        if t.startswith("AS:i:"):
            AS = int(t[5:])
        elif t.startswith("XS:i:"):
            XS = int(t[5:])
    return Alignment(f[0], int(f[1]), name2tid.get(f[2], -1), int(f[3]) - 1, cigar, None if f[9] == "*" else f[9],
                     None if f[10] == "*" else f[10], AS, XS, line_no)


def read_sam(fh) -> Tuple[List[str], List[int], Iterator[Alignment]]:
    names: List[str] = []
    lengths: List[int] = []
    name2tid = {}
    first = None
    for line in fh:
        if line.startswith("@"):
            if line.startswith("@SQ"):
                sn, ln = None, 0
                for x in line.rstrip("\n").split("\t")[1:]:
                    if x.startswith("SN:"):
                        sn = x[3:]
                    elif x.startswith("LN:"):
                        ln = int(x[3:])
                name2tid[sn] = len(names)
                names.append(sn)
                lengths.append(ln)
            continue
        first = line
        break

    def it():
        n = 0
        if first is not None and first.strip():
            yield _parse_sam_line(first, name2tid, n)
            n += 1
        for line in fh:
            if not line.strip():
                continue
            yield _parse_sam_line(line, name2tid, n)
            n += 1

    return names, lengths, it()


_SEQ_CODE = "=ACMGRSVTWYHKDBN"
_CIG_CODE = "MIDNSHP=X"


def read_bam(fh) -> Tuple[List[str], List[int], Iterator[Alignment]]:
    """fh: binary file object of a BAM file (BGZF).  Only the fields the hot path needs are decoded."""
    z = gzip.GzipFile(fileobj=fh, mode="rb")

    def need(n):
        b = z.read(n)
        if len(b) != n:
            raise EOFError
        return b

    if need(4) != b"BAM\x01":
        raise IOError("not a BAM file")
    (l_text,) = struct.unpack("<i", need(4))
    need(l_text)
    (n_ref,) = struct.unpack("<i", need(4))
    names, lengths = [], []
    for _ in range(n_ref):
        (l_name,) = struct.unpack("<i", need(4))
        names.append(need(l_name)[:-1].decode("latin-1"))
        lengths.append(struct.unpack("<i", need(4))[0])

    def it():
        n = 0
        while True:
            head = z.read(4)
            if len(head) < 4:
                return
            (block,) = struct.unpack("<i", head)
            b = need(block)
            tid, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _nt, _np, _tl = struct.unpack("<iiBBHHHiiii", b[:32])
            o = 32
            qname = b[o : o + l_rn - 1].decode("latin-1")
            o += l_rn
            cig = []
            for k in range(n_cig):
                (v,) = struct.unpack_from("<I", b, o + 4 * k)
                cig.append((v & 0xF, v >> 4))
            o += 4 * n_cig
            nb = (l_seq + 1) // 2
            sb = b[o : o + nb]
            o += nb
            seq = "".join(_SEQ_CODE[x >> 4] + _SEQ_CODE[x & 15] for x in sb)[:l_seq] if l_seq else None
            qb = b[o : o + l_seq]
            o += l_seq
            qual = None if (l_seq == 0 or qb[0] == 0xFF) else "".join(chr(x + 33) for x in qb)
            AS = XS = None
            while o < len(b):
                tag = b[o : o + 2]
                typ = chr(b[o + 2])
                o += 3
                if typ in "cCsSiI":
                    fmt, sz = {"c": ("<b", 1), "C": ("<B", 1), "s": ("<h", 2), "S": ("<H", 2), "i": ("<i", 4), "I": ("<I", 4)}[typ]
                    (v,) = struct.unpack_from(fmt, b, o)
                    o += sz
                    if tag == b"AS":
                        AS = v
                    elif tag == b"XS":
                        XS = v
                elif typ == "A":
                    o += 1
                elif typ == "f":
                    o += 4
                elif typ in "ZH":
                    e = b.index(b"\x00", o)
                    o = e + 1
                elif typ == "B":
                    sub = chr(b[o])
                    (cnt,) = struct.unpack_from("<i", b, o + 1)
                    o += 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
                else:
                    break
            yield Alignment(qname, flag, tid, pos, cig or None, seq, qual, AS, XS, n)
            n += 1

    return names, lengths, it()


def open_alignments(path: Optional[str]):
    """the reference's rule (find_circ.py:461-469): names ending in 'sam' are SAM text, anything else BAM; no
    argument = SAM text on stdin"""
    if path is None or path == "-":
        return read_sam(sys.stdin)
    if path.endswith("sam"):
        return read_sam(open(path, encoding="latin-1"))  # bytes as they are, like the native ingest
    return read_bam(open(path, "rb"))


def sam_header_names(path: str) -> List[str]:
    """@SQ names of a SAM text file (the native ingest parses the body itself)"""
    names = []
    with open(path, "rb") as fh:
        for line in fh:
            if not line.startswith(b"@"):
                break
            if line.startswith(b"@SQ"):
                for x in line.rstrip(b"\r\n").split(b"\t")[1:]:
                    if x.startswith(b"SN:"):
                        names.append(x[3:].decode("latin-1"))
    return names
