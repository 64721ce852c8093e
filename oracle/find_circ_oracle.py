"""
CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

An own-words python3 restatement of the hot path of the reference (/root/reference/find_circ.py, v1.99):
fragment grouping -> adjacent segment pairs -> breakpoint scan -> per-fragment evidence -> junction
aggregation -> BED / reads / multi-event / counter output.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the product
(find_circ2_b200/) never does and fails loudly when its CUDA library is missing.

Parity status: PINNED.  tests/test_oracle_golden.py checks this module byte-for-byte (after the canonical
row sort that BASELINE.json allows) against tests/golden/*/ref_*/, which were produced by the reference's
own find_circ.py executed through oracle/ref_shim/run_reference.py (python2->python3 mechanics only), on
the reference's test_data fixtures and on seeded synthetic inputs, under 13 option sets.

Each function cites the reference lines it follows.  Deliberately mirrored quirks:
  * python2 str(float) (12 significant digits) in the BED columns n_weight / uniq_bridges / tiss_counts
    and in the counters (find_circ.py:597, 724-730, 1607);
  * with --max-mismatch 0 the edit distance is a bool and prints as 'False' (find_circ.py:868-870, 727);
  * the first SAM record is never checked for being unmapped (find_circ.py:1462-1463);
  * counters anchor_not_uniq / no_uniq_bridges are incremented after the dump and never appear (:1605-1610).
Not supported (the reference itself crashes or needs absent libraries): --stranded (find_circ.py:533 vs
:766-776), -S/--system, -B/--bam.
"""
from __future__ import annotations

import dataclasses
import re
from collections import defaultdict
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------------
# options (find_circ.py:383-413)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Options:
    name: str = "unknown"  # -n
    min_uniq_qual: int = 2  # --min-uniq-qual
    asize: int = 15  # -a
    margin: int = 2  # -m
    maxdist: int = 2  # -d
    short_threshold: int = 100
    huge_threshold: int = 100000
    noncanonical: bool = False
    allhits: bool = False
    strandpref: bool = False
    halfunique: bool = False
    report_nobridges: bool = False
    nolinear: bool = False
    multi_events: bool = True
    known_circ: str = ""
    known_lin: str = ""
    test: bool = False   # compare every fragment with the truth encoded in its name (find_circ.py:411, 1148-1273)
    noop: bool = False   # only group the alignments (find_circ.py:1554-1558)
    stdout: str = ""     # circs | lins | reads | multi: that output goes to stdout (find_circ.py:453-458)


def py2_str(x) -> str:
    """python2's str(): floats carry 12 significant digits and always a fractional part."""
    if isinstance(x, (bool, np.bool_)):
        return "True" if x else "False"
    if isinstance(x, float):
        s = "%.12g" % x
        if s in ("inf", "-inf", "nan"):
            return s
        if "." not in s and "e" not in s:
            s += ".0"
        return s
    return str(x)


_COMPLEMENT = {
    "a": "t", "t": "a", "c": "g", "g": "c", "k": "m", "m": "k", "r": "y", "y": "r", "s": "s", "w": "w",
    "b": "v", "v": "b", "h": "d", "d": "h", "n": "n",
}  # find_circ.py:21-52
_COMPLEMENT.update({k.upper(): v.upper() for k, v in list(_COMPLEMENT.items())})


def rev_comp(seq: str) -> str:
    """find_circ.py:54-58 (KeyError on letters outside the IUPAC table, like the reference)"""
    return "".join(_COMPLEMENT[c] for c in reversed(seq))


# ----------------------------------------------------------------------------------------------
# genome access (find_circ.py:103-215, 329-371)
# ----------------------------------------------------------------------------------------------
class Genome:
    """Whole FASTA in memory; `get` has the reference's semantics for windows that overlap the chromosome:
    positions outside [0,size) read as 'N' (find_circ.py:194-211), case is preserved (callers upper-case).
    Windows lying ENTIRELY outside the chromosome are undefined in the reference (wrong-length strings,
    find_circ.py:196-209); here they read as all-N."""

    def __init__(self, path: str):
        self.names: List[str] = []
        self.seqs: Dict[str, str] = {}
        name, chunks = None, []
        with open(path) as fh:
            for line in fh:
                if line.startswith(">"):
                    if name is not None:
                        self.seqs[name] = "".join(chunks)
                    name = line[1:].split()[0].strip()  # find_circ.py:136
                    self.names.append(name)
                    chunks = []
                else:
                    chunks.append(line.strip())
        if name is not None:
            self.seqs[name] = "".join(chunks)

    def size(self, chrom: str) -> int:
        return len(self.seqs[chrom])

    def get(self, chrom: str, start: int, end: int) -> str:
        seq = self.seqs[chrom]  # KeyError for unknown chromosomes, like find_circ.py:193
        size = len(seq)
        if end <= start:
            return ""
        left = max(0, -start) if start < 0 else 0
        right = max(0, end - size) if end > size else 0
        left = min(left, end - start)
        right = min(right, end - start)
        core = seq[max(start, 0) : max(min(end, size), 0)]
        return "N" * left + core + "N" * right


# ----------------------------------------------------------------------------------------------
# SAM text records with pysam's field semantics
# ----------------------------------------------------------------------------------------------
_CIG = re.compile(r"(\d+)([MIDNSHP=X])")
_OPCODE = {c: i for i, c in enumerate("MIDNSHP=X")}


class Record:
    __slots__ = ("qname", "flag", "tid", "pos", "cigar", "seq", "qual", "AS", "XS", "line_no")

    def __init__(self, fields: List[str], tid: int, line_no: int):
        self.qname = fields[0]
        self.flag = int(fields[1])
        self.tid = tid
        self.pos = int(fields[3]) - 1
        self.cigar = None if fields[5] == "*" else [(_OPCODE[c], int(n)) for n, c in _CIG.findall(fields[5])]
        self.seq = None if fields[9] == "*" else fields[9]
        self.qual = None if fields[10] == "*" else fields[10]
        self.AS = None
        self.XS = None
        for t in fields[11:]:
            if t.startswith("AS:i:"):
                self.AS = int(t[5:])
            elif t.startswith("XS:i:"):
                self.XS = int(t[5:])
        self.line_no = line_no

    is_unmapped = property(lambda s: bool(s.flag & 0x4))
    is_reverse = property(lambda s: bool(s.flag & 0x10))
    is_read1 = property(lambda s: bool(s.flag & 0x40))
    is_read2 = property(lambda s: bool(s.flag & 0x80))

    @property
    def aend(self) -> Optional[int]:
        if self.is_unmapped or not self.cigar:
            return None
        return self.pos + sum(n for op, n in self.cigar if op in (0, 2, 3, 7, 8))

    @property
    def query_len(self) -> int:
        """len(pysam .query): the sequence without soft clips (hard clips are not stored in SEQ)."""
        n = len(self.seq)
        if self.cigar:
            for seq_ in (self.cigar, reversed(self.cigar)):
                for op, c in seq_:
                    if op == 5:
                        continue
                    if op == 4:
                        n -= c
                    else:
                        break
        return n

    @property
    def read_start(self) -> int:
        """offset of the aligned part inside the full read (find_circ.py:1086-1097): clip lengths are
        summed until the first M; other operations are skipped without ending the loop."""
        start = 0
        for op, c in self.cigar:
            if op in (4, 5):
                start += c
            elif op == 0:
                break
        return start

    def uniqueness(self) -> int:
        """AS - XS, XS absent -> AS (find_circ.py:809-819)"""
        return self.AS - self.XS if self.XS is not None else self.AS


def read_sam(path_or_lines) -> Tuple[List[str], Iterator[Record]]:
    """returns (@SQ names, iterator over alignment records)"""
    fh = open(path_or_lines) if isinstance(path_or_lines, str) else iter(path_or_lines)
    names: List[str] = []
    name2tid: Dict[str, int] = {}
    first_body = None
    for line in fh:
        if line.startswith("@"):
            if line.startswith("@SQ"):
                for f in line.rstrip("\n").split("\t")[1:]:
                    if f.startswith("SN:"):
                        name2tid[f[3:]] = len(names)
                        names.append(f[3:])
            continue
        first_body = line
        break

    def records() -> Iterator[Record]:
        n = 0
        if first_body is not None:
            f = first_body.rstrip("\r\n").split("\t")
            yield Record(f, name2tid.get(f[2], -1), n)
            n += 1
        for line in fh:
            if not line.strip():
                continue
            f = line.rstrip("\r\n").split("\t")
            yield Record(f, name2tid.get(f[2], -1), n)
            n += 1

    return names, records()


# ----------------------------------------------------------------------------------------------
# fragments, mates, spans (find_circ.py:976-1140, 821-852, 1450-1486)
# ----------------------------------------------------------------------------------------------
class Mate:
    def __init__(self, primary: Record, counters):
        counters["total_mates"] += 1  # find_circ.py:978
        self.primary = primary
        self.proper: List[Record] = [primary]
        self.other_chrom: List[Record] = []
        self.other_strand: List[Record] = []

    def add(self, rec: Record) -> None:
        """find_circ.py:1039-1056"""
        if rec.tid != self.primary.tid:
            self.other_chrom.append(rec)
        elif rec.is_reverse != self.primary.is_reverse:
            self.other_strand.append(rec)
        else:
            self.proper.append(rec)


@dataclasses.dataclass(eq=False)
class Span:
    """one anchor pair (JunctionSpan, find_circ.py:821-852)"""

    A: Record
    B: Record
    primary: Record
    q_start: int
    q_end: int
    weight: float

    def __post_init__(self):
        self.uniq = min(self.A.uniqueness(), self.B.uniqueness())
        self.strand = "-" if self.primary.is_reverse else "+"
        self.dist = self.B.pos - self.A.aend
        self.read_part = self.primary.seq[self.q_start : self.q_end]

    @property
    def is_backsplice(self) -> bool:
        return self.dist < 0


def fragments(records: Iterable[Record], counters) -> Iterator[Tuple[Optional[Mate], Mate]]:
    """find_circ.py:1450-1486: consecutive records with one name form a fragment; a change of the read1
    flag switches to the other mate; unmapped records (except the very first) are counted and skipped."""
    it = iter(records)
    try:
        first = next(it)
    except StopIteration:
        return
    cur, other = Mate(first, counters), None
    for rec in it:
        if rec.is_unmapped:
            counters["unmapped_reads"] += 1
            continue
        same_name = rec.qname == cur.primary.qname
        if same_name and rec.is_read1 == cur.primary.is_read1:
            cur.add(rec)
        elif same_name:
            other, cur = cur, Mate(rec, counters)
        else:
            yield other, cur
            other, cur = None, Mate(rec, counters)
    yield other, cur


def adjacent_spans(mate: Mate, opt: Options, counters) -> Iterator[Span]:
    """find_circ.py:1058-1140"""
    segs = mate.proper
    if len(segs) < 2:
        return
    weight = 1.0 / (len(segs) - 1.0)
    start = {id(s): s.read_start for s in segs}
    end = {id(s): start[id(s)] + s.query_len for s in segs}
    in_read_order = sorted(segs, key=lambda s: start[id(s)])
    for a, b in zip(in_read_order, in_read_order[1:]):
        la = end[id(a)] - start[id(a)]
        lb = end[id(b)] - start[id(b)]
        if la < opt.asize or lb < opt.asize:
            counters["seg_too_short_skip"] += 1
            continue
        yield Span(a, b, mate.primary, min(start[id(a)], start[id(b)]), max(end[id(a)], end[id(b)]), weight)


# ----------------------------------------------------------------------------------------------
# breakpoint scan (find_circ.py:854-974, 766-806)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass(eq=False)
class Splice:
    span: Optional[Span]
    chrom: str
    start: int
    end: int
    strand: str
    dist: object  # int, or bool when --max-mismatch 0
    ov: int
    gtag: str
    n_hits: int = 1
    span_strand: str = "+"

    @property
    def coord(self):
        lo, hi = (self.start, self.end) if self.start < self.end else (self.end, self.start)
        return (self.chrom, lo, hi, self.strand)

    def score(self, strandpref: bool) -> int:
        """find_circ.py:792-799"""
        s = (self.gtag == "GTAG") * 20 - self.dist * 10 - self.ov
        if strandpref:
            s += 100 * (self.strand == self.span_strand)
        return s


_IDX = {c: i for i, c in enumerate("ACGTN")}


def _rc4(mer: str) -> str:
    """find_circ.py:70-74: only {A,C,G,T,N}^4 is tabulated; anything else is a KeyError there"""
    for c in mer:
        if c not in _IDX:
            raise KeyError(mer)
    return "".join("TGCAN"[_IDX[c]] for c in reversed(mer))


def _mismatches(a: str, b: str):
    """find_circ.py:861-863 -- one numpy byte comparison per split position (the reference's cost model)"""
    x = np.frombuffer(a.encode("latin-1"), dtype=np.byte)
    y = np.frombuffer(b.encode("latin-1"), dtype=np.byte)
    return (x != y).sum()


def scan_windows(
    a_flank: str,
    b_flank: str,
    internal: str,
    chrom: str,
    a_start: int,
    b_end: int,
    is_backsplice: bool,
    span_strand: str,
    opt: Options,
    span: Optional[Span] = None,
) -> List[Splice]:
    """The loop of find_circ.py:904-974 on already-fetched windows.
    a_start = A.pos + eff (genomic position of a_flank[0]); b_end = B.aend - eff (one past b_flank[-1])."""
    l = len(internal)
    hits: List[Splice] = []
    exact_only = opt.maxdist == 0
    for x in range(l + 1):
        spliced = a_flank[:x] + b_flank[x + 2 :]
        dist = (spliced != internal) if exact_only else _mismatches(spliced, internal)
        if not dist <= opt.maxdist:
            continue
        ov = 0
        if opt.margin:
            if x < opt.margin:
                ov = opt.margin - x
            if l - x < opt.margin:
                ov = opt.margin - (l - x)
        gtag = a_flank[x : x + 2] + b_flank[x : x + 2]
        rc = _rc4(gtag)
        start, end = b_end - l + x, a_start + x + 1
        start, end = min(start, end), max(start, end)
        if is_backsplice:
            end -= 1
        else:
            start -= 1
        if not exact_only:
            dist = int(dist)
        if opt.noncanonical:
            hits.append(Splice(span, chrom, start, end, "+", dist, ov, gtag, 1, span_strand))
            hits.append(Splice(span, chrom, start, end, "-", dist, ov, rc, 1, span_strand))
        elif gtag == "GTAG":
            hits.append(Splice(span, chrom, start, end, "+", dist, ov, gtag, 1, span_strand))
        elif gtag == "CTAC":
            hits.append(Splice(span, chrom, start, end, "-", dist, ov, rc, 1, span_strand))
    if len(hits) < 2:
        return hits
    ranked = sorted(hits, key=lambda h: h.score(opt.strandpref), reverse=True)  # stable: ties keep x order
    best = ranked[0].score(opt.strandpref)
    ties = [h for h in ranked if h.score(opt.strandpref) == best]
    for h in ranked:
        h.n_hits = len(ties)
    return ties


def window_geometry(a_pos: int, b_aend: int, L: int, opt: Options) -> Tuple[int, int, int]:
    """(a_start, b_end, l) of find_circ.py:882, 900-904"""
    eff = opt.asize - opt.margin
    return a_pos + eff, b_aend - eff, L - 2 * eff


def find_breakpoints(span: Span, genome: Genome, chrom: str, opt: Options) -> List[Splice]:
    """find_circ.py:854-974 for one span"""
    eff = opt.asize - opt.margin
    L = len(span.read_part)
    a_start, b_end, l = window_geometry(span.A.pos, span.B.aend, L, opt)
    # read[eff:-eff] (find_circ.py:895); python clamps a negative end that reaches past the start
    internal = span.read_part[eff : max(L - eff, 0)].upper() if eff > 0 else ""
    flank = l + 2
    a_flank = genome.get(chrom, a_start, a_start + flank).upper()
    b_flank = genome.get(chrom, b_end - flank, b_end).upper()
    if l < 0:
        return []
    return scan_windows(a_flank, b_flank, internal, chrom, a_start, b_end, span.is_backsplice, span.strand, opt, span)


# ----------------------------------------------------------------------------------------------
# junction aggregation (find_circ.py:486-730)
# ----------------------------------------------------------------------------------------------
class Junction:
    """accumulator for one (chrom,start,end,strand) (Hit, find_circ.py:486-654)"""

    def __init__(self, name: str, coord):
        self.name = name
        self.coord = coord
        self.n_spanned = 0
        self.n_weighted = 0.0
        self.n_uniq_bridges = 0.0
        self.edits: List = []
        self.overlaps: List[int] = []
        self.n_hits: List[int] = []
        self.quals_left: List[int] = []
        self.quals_right: List[int] = []
        self.readnames: List[str] = []
        self.uniq = set()
        self.signal = "NNNN"
        self.tissues: Dict[str, float] = defaultdict(float)
        self.flags: Dict[str, int] = defaultdict(int)
        self.read_flags: Dict[str, set] = defaultdict(set)

    def add(self, sp: Splice, sample: str) -> None:
        """find_circ.py:526-582"""
        self.signal = sp.gtag
        self.edits.append(sp.dist)
        self.overlaps.append(sp.ov)
        self.n_hits.append(sp.n_hits)
        span = sp.span
        self.n_spanned += 1
        self.n_weighted += span.weight
        left, right = (span.B, span.A) if span.is_backsplice else (span.A, span.B)
        q_left = left.AS - (left.XS if left.XS is not None else 0)
        q_right = right.AS - (right.XS if right.XS is not None else 0)
        if q_left and q_right:
            self.n_uniq_bridges += span.weight
        self.quals_left.append(q_left)
        self.quals_right.append(q_right)
        self.readnames.append(span.primary.qname)
        read = span.primary.seq
        self.tissues[sample] += span.weight
        self.uniq.add((read, sample))
        self.uniq.add((rev_comp(read), sample))

    def add_flag(self, flag: str, frag: str) -> None:
        self.flags[flag] += 1
        self.read_flags[frag].add(flag)

    def categories(self, opt: Options) -> List[str]:
        """find_circ.py:601-654"""
        cats = []
        if self.signal != "GTAG":
            cats.append("NON_CANONICAL")
        if max(self.quals_left) == 0 or max(self.quals_right) == 0:
            cats.append("WARN_NON_UNIQUE_ANCHOR")
        if self.n_uniq_bridges == 0:
            cats.append("WARN_NO_UNIQ_BRIDGES")
        if min(self.n_hits) > 1:
            cats.append("WARN_AMBIGUOUS_BP")
        ov, ed = min(self.overlaps), min(self.edits)
        if ov == 0 and ed == 0:
            pass
        elif ov < 2 and ed < 2:
            cats.append("WARN_EXT_1MM")
        elif ov >= 2 or ed >= 2:
            cats.append("WARN_EXT_2MM+")
        _, start, end, _ = self.coord
        if end - start < opt.short_threshold:
            cats.append("SHORT")
        elif end - start > opt.huge_threshold:
            cats.append("HUGE")
        if self.read_flags:
            unbroken = sum(1 for fl in self.read_flags.values() if "BROKEN_SEGMENTS" not in fl)
            # counts non-WARN *flags*, not reads (find_circ.py:640-642)
            unwarned = sum(1 for fl in self.read_flags.values() for w in fl if not w.startswith("WARN"))
            if not unbroken:
                cats.append("WARN_ALWAYS_BROKEN")
            if not unwarned:
                cats.append("WARN_ALWAYS_WARN")
        return cats


BED_HEADER = [
    "chrom", "start", "end", "name", "n_frags", "strand", "n_weight", "n_spanned", "n_uniq", "uniq_bridges",
    "best_qual_left", "best_qual_right", "tissues", "tiss_counts", "edits", "anchor_overlap", "breakpoints",
    "signal", "strandmatch", "category", "flags", "flag_counts",
]  # find_circ.py:696


class JunctionTable:
    """SpliceSiteStorage, find_circ.py:657-730"""

    def __init__(self, prefix: str, opt: Options, known: str = ""):
        self.prefix = prefix
        self.opt = opt
        self.sites: Dict[tuple, Junction] = {}
        self.count = 0
        if known:
            self.load_known(known)

    def load_known(self, path: str) -> None:
        """find_circ.py:665-679: BED6 sites are pre-seeded under their own names with a placeholder splice whose
        dist=10, ov=10, n_hits=1 stay in the per-junction lists (and therefore in the minima that are printed)"""
        for line in open(path):
            if line.startswith("#"):
                continue
            parts = line.rstrip().split("\t")
            chrom, start, end, name, _score, strand = parts[:6]
            coord = (chrom, int(start), int(end), strand)
            j = Junction(name, coord)
            j.edits.append(10)
            j.overlaps.append(10)
            j.n_hits.append(1)
            self.sites[coord] = j

    def add(self, sp: Splice) -> Junction:
        coord = sp.coord
        j = self.sites.get(coord)
        if j is None:
            self.count += 1
            j = Junction("%s_%s_%06d" % (self.opt.name, self.prefix, self.count), coord)
            self.sites[coord] = j
        j.add(sp, self.opt.name)
        return j

    def rows(self) -> List[str]:
        """find_circ.py:695-730 (without the header); discovery order"""
        opt = self.opt
        out = []
        for (chrom, start, end, strand), j in self.sites.items():
            if not j.readnames:
                continue  # a known site that was not observed (find_circ.py:699-700)
            ql, qr = max(j.quals_left), max(j.quals_right)
            if opt.halfunique:
                if ql < opt.min_uniq_qual and qr < opt.min_uniq_qual:
                    continue
            elif ql < opt.min_uniq_qual or qr < opt.min_uniq_qual:
                continue
            if j.n_uniq_bridges == 0 and not opt.report_nobridges:
                continue
            tissues = sorted(j.tissues)
            counts = [py2_str(j.tissues[t]) for t in tissues]
            if j.flags:
                flags = sorted(j.flags)
                fcounts = [j.flags[f] for f in flags]
            else:
                flags, fcounts = ["N/A"], [0]
            cols = [
                chrom, start, end, j.name, len(set(j.readnames)), strand,
                j.n_weighted, j.n_spanned, len(j.uniq) // 2, j.n_uniq_bridges,
                ql, qr, ",".join(tissues), ",".join(counts),
                min(j.edits), min(j.overlaps), min(j.n_hits), j.signal, "N/A",
                ",".join(sorted(j.categories(opt))), ",".join(flags), ",".join(str(c) for c in fcounts),
            ]
            out.append("\t".join(py2_str(c) for c in cols))
        return out

    def text(self) -> str:
        return "#" + "\t".join(BED_HEADER) + "\n" + "".join(r + "\n" for r in self.rows())


MULTI_HEADER = ["chrom", "start", "end", "name", "score", "strand", "fragment_name", "lin_cons", "lin_incons",
                "unspliced_cons", "unspliced_incons"]  # find_circ.py:735


def multi_event_row(frag: str, circ: Junction, lin_cons, lin_incons, un_cons, un_incons) -> str:
    """find_circ.py:737-763 (sets are iterated in python's order there; sorted here, compared sorted)"""
    score = len(lin_cons) - 10 * len(lin_incons) + len(un_cons) - 10 * len(un_incons)
    chrom, start, end, strand = circ.coord
    cols = [chrom, str(start), str(end), "ME:" + circ.name, str(score), strand, frag]
    cols.append(",".join("%d-%d" % (s, e) for c, s, e, _ in sorted(lin_cons)) if lin_cons else "NO_LIN_CONS")
    cols.append(",".join("[%s:%d-%d]" % (c, s, e) for c, s, e, _ in sorted(lin_incons)) if lin_incons else "NO_LIN_INCONS")
    cols.append(",".join("%d-%d" % (s, e) for c, s, e, _ in sorted(un_cons)) if un_cons else "NO_UNSPLICED_CONS")
    cols.append(
        ",".join("[%s:%d-%d]" % (c, s, e) for c, s, e, _ in sorted(un_incons)) if un_incons else "NO_UNSPLICED_INCONS"
    )
    return "\t".join(cols)


# ----------------------------------------------------------------------------------------------
# the run (find_circ.py:1276-1439, 1490-1610)
# ----------------------------------------------------------------------------------------------
def truth_from_name(text: str):
    """the fragment structure a simulated read carries behind '___' in its name (find_circ.py:1147-1191):
    mates are separated by '|', steps by ';'.  O:chrom:start:strand sets the origin, M:n walks n bases, LS:a:b / CS:a:b are
    a linear / circular splice with both ends relative to the origin.  Returns (linear, circular, unspliced) coordinate sets;
    origin and position carry over from one mate to the next, as they do in the reference."""
    chrom = strand = None
    start = end = None
    lin, circ, unspliced = set(), set(), set()
    for mate in text.split("|"):
        spliced = False
        for step in mate.split(";"):
            f = step.split(":")
            if f[0] == "O":
                chrom, start, strand = f[1], int(f[2]), f[3]
                end = start
            elif f[0] == "M":
                end += int(f[1])
            elif f[0] == "LS":
                left, right = int(f[1]) + start, int(f[2]) + start
                lin.add((chrom, left, right, strand))
                spliced, end = True, right
            elif f[0] == "CS":
                left, right = int(f[1]) + start, int(f[2]) + start
                circ.add((chrom, left, right, strand))
                spliced, end = True, left
        if not spliced and chrom:
            unspliced.add((chrom, start, end, "*"))  # (strand only with --stranded, which the reference cannot run)
    return lin, circ, unspliced


def test_row(frag: str, lin_coords, circ_coords, unspliced_coords, broken_coords) -> str:
    """one line of test_results.tsv (find_circ.py:1194-1273)"""
    if "___" not in frag:
        return "\t".join([frag, "N/A", "N/A", "N/A", "N/A"])
    lin_ref, circ_ref, un_ref = truth_from_name(frag.split("___")[-1])
    listing = lambda coords: ",".join(str(c) for c in sorted(coords))  # noqa: E731

    def verdict(ref, got, what, ok):
        flags = set()
        if ref - got:
            flags.add("MISSED_%s:%s" % (what, listing(ref - got)))
        if got - ref:
            flags.add("SPURIOUS_%s:%s" % (what, listing(got - ref)))
        if flags:
            return ";".join(sorted(flags))
        return ok if ref else "N/A"

    cols = [frag, verdict(lin_ref, set(lin_coords), "LINEAR_JUNCTIONS", "LIN_OK"),
            verdict(circ_ref, set(circ_coords), "CIRCULAR_JUNCTIONS", "CIRC_OK"),
            verdict(un_ref, set(unspliced_coords), "UNSPLICED", "UNSPLICED_OK"),
            "BROKEN_SEGMENTS:" + ";".join(str(b) for b in sorted(broken_coords)) if broken_coords else "N/A"]
    return "\t".join(cols)


@dataclasses.dataclass
class Outputs:
    circ_bed: str
    lin_bed: str
    reads_fastq: str
    multi_events: str
    counters: str  # 'key=value' lines as dumped to run.log
    test_results: str = ""  # --test
    n_fragments: int = 0
    n_spans: int = 0


class Run:
    def __init__(self, genome: Genome, chrom_names: Sequence[str], opt: Options):
        self.genome = genome
        self.chrom_names = list(chrom_names)
        self.opt = opt
        self.N: Dict[str, float] = defaultdict(float)
        self.circ = JunctionTable("circ", opt, opt.known_circ)
        self.lin = JunctionTable("lin", opt, opt.known_lin)
        self.reads: List[str] = []
        self.multi: List[str] = []
        self.test_rows: List[str] = []
        self.n_fragments = 0
        self.n_spans = 0

    # -- find_circ.py:1276-1439 ------------------------------------------------------------
    def record_hits(self, frag: str, circ_spans: List[Span], lin_spans: List[Span], unspliced: List[Record],
                    broken: List[Record]):
        opt, N = self.opt, self.N
        warns, junctions = set(), []

        def note(j):
            if not any(j is k for k in junctions):
                junctions.append(j)

        circ_coords = set()
        circ = None
        for span in circ_spans:
            if span.uniq < opt.min_uniq_qual:
                N["circ_junc_not_unique"] += 1
                continue
            self.n_spans += 1
            splices = find_breakpoints(span, self.genome, self.chrom_names[span.A.tid], opt)
            if not splices:
                N["circ_no_bp"] += 1
                warns.add("WARN_UNRESOLVED_EXTRA_BACKSPLICE")
                continue
            N["circ_spliced"] += 1
            for sp in splices:
                circ = self.circ.add(sp)
                circ_coords.add(circ.coord)
                note(circ)
                if not opt.allhits:
                    break

        if len(circ_coords) > 1:
            for coord in circ_coords:
                warns.add("WARN_MULTI_BACKSPLICE")
                c = self.circ.sites[coord]
                c.add_flag("WARN_MULTI_BACKSPLICE", frag)
                note(c)
            return junctions, warns
        if not circ_coords and opt.nolinear:
            return junctions, warns

        if circ_coords:
            _, circ_start, circ_end, _ = circ.coord
            first_span = circ_spans[0]
            if len(circ_spans) > 1:
                warns.add("SUPPORT_CLOSURE")

        lin_cons, lin_incons = set(), set()
        lin_coords = set()
        for span in lin_spans:
            if span.uniq < opt.min_uniq_qual:
                N["lin_junc_not_unique"] += 1
                continue
            self.n_spans += 1
            splices = find_breakpoints(span, self.genome, self.chrom_names[span.A.tid], opt)
            if not splices:
                N["lin_no_bp"] += 1
                warns.add("WARN_UNRESOLVED_LINSPLICE")
                continue
            N["lin_spliced"] += 1
            for sp in splices:
                lin = self.lin.add(sp)
                note(lin)
                lin_coords.add(lin.coord)
                if circ_coords:
                    if sp.start <= circ_start or sp.end >= circ_end:
                        warns.add("WARN_OUTSIDE_SPLICE_JUNCTION")
                        lin_incons.add(sp.coord)
                    else:
                        lin_cons.add(sp.coord)
                        warns.add("SUPPORT_INSIDE_SPLICE_JUNCTION")
                if not opt.allhits:
                    break

        if opt.test:  # find_circ.py:1380-1394
            star = lambda rec: (self.chrom_names[rec.tid], rec.pos, rec.aend, "*")  # noqa: E731
            self.test_rows.append(test_row(frag, lin_coords, circ_coords, {star(r) for r in unspliced}, {star(r) for r in broken}))

        if circ_coords:
            un_cons, un_incons = set(), set()
            for rec in unspliced:
                coord = (self.chrom_names[rec.tid], rec.pos, rec.aend, "*")
                if first_span.primary.tid != rec.tid:
                    warns.add("WARN_OTHER_CHROM_MATE")
                    un_incons.add(coord)
                elif rec.pos + opt.asize <= circ_start or rec.aend - opt.asize >= circ_end:
                    warns.add("WARN_OUTSIDE_MATE")
                    un_incons.add(coord)
                else:
                    warns.add("SUPPORT_INSIDE_MATE")
                    un_cons.add(coord)
            if broken:
                warns.add("BROKEN_SEGMENTS")
            if (un_cons or un_incons or lin_cons or lin_incons) and opt.multi_events:
                self.multi.append(multi_event_row(frag, circ, lin_cons, lin_incons, un_cons, un_incons))
            for w in warns:
                circ.add_flag(w, frag)
        return junctions, warns

    # -- find_circ.py:1490-1574 ------------------------------------------------------------
    def process(self, records: Iterable[Record]) -> None:
        opt, N = self.opt, self.N
        for mate1, mate2 in fragments(records, N):
            self.n_fragments += 1
            if opt.noop:
                continue
            frag = mate2.primary.qname
            circ_spans, lin_spans, unspliced, broken = [], [], [], []
            for mate in (mate1, mate2):
                if mate is None:
                    continue
                if len(mate.proper) < 2:
                    N["unspliced_mates"] += 1
                    unspliced.append(mate.primary)
                    continue
                L = len(mate.primary.seq)
                lo, hi = L, 0
                for span in adjacent_spans(mate, opt, N):
                    (circ_spans if span.is_backsplice else lin_spans).append(span)
                    lo, hi = min(lo, span.q_start), max(hi, span.q_end)
                if hi < L - opt.asize or lo > opt.asize:
                    broken.extend(mate.other_chrom)
                    broken.extend(mate.other_strand)
            if not circ_spans and opt.nolinear:
                continue
            if circ_spans or lin_spans:
                junctions, flags = self.record_hits(frag, circ_spans, lin_spans, unspliced, broken)
                if junctions:
                    for mate in (mate1, mate2):
                        if mate is not None:
                            self.reads.append(format_read(mate.primary, [j.name for j in junctions], flags))

    def outputs(self) -> Outputs:
        counters = "".join("%s=%s\n" % (k, py2_str(self.N[k])) for k in sorted(self.N))
        return Outputs(
            circ_bed=self.circ.text(),
            lin_bed=self.lin.text(),
            reads_fastq="".join(self.reads),
            multi_events="#" + "\t".join(MULTI_HEADER) + "\n" + "".join(r + "\n" for r in self.multi),
            counters=counters,
            test_results="".join(r + "\n" for r in self.test_rows),
            n_fragments=self.n_fragments,
            n_spans=self.n_spans,
        )


def format_read(primary: Record, junction_names: Sequence[str], flags: Iterable[str]) -> str:
    """find_circ.py:1442-1447"""
    name = "%s %s %s" % (primary.qname, ",".join(sorted(junction_names)), ",".join(sorted(flags)))
    return "@%s\n%s\n+%s\n%s\n" % (name, primary.seq, name, primary.qual)


def run(genome_path: str, sam_path: str, opt: Options) -> Outputs:
    genome = Genome(genome_path)
    names, records = read_sam(sam_path)
    r = Run(genome, names, opt)
    r.process(records)
    return r.outputs()


# ----------------------------------------------------------------------------------------------
# helpers for comparisons
# ----------------------------------------------------------------------------------------------
def canonical_bed(text: str) -> List[str]:
    """header first, then rows sorted (python2 dict order is not reproducible; BASELINE.json allows this)"""
    lines = [l for l in text.split("\n") if l]
    head = [l for l in lines if l.startswith("#")]
    body = sorted(l for l in lines if not l.startswith("#"))
    return head + body


def canonical_multi(text: str) -> List[str]:
    """multi_events rows: the comma lists inside columns 8-11 come from python sets -> sort them too"""
    out = []
    for l in text.split("\n"):
        if not l:
            continue
        if l.startswith("#"):
            out.append(l)
            continue
        c = l.split("\t")
        for k in range(7, 11):
            c[k] = ",".join(sorted(c[k].split(",")))
        out.append("\t".join(c))
    return out[:1] + sorted(out[1:])


def options_from_argv(argv: Sequence[str]) -> Options:
    """parse the shipped command-line spellings (find_circ.py:383-413) into Options (tests use this to
    replay tests/golden/*/ref_*/cmdline.txt)"""
    o = Options()
    it = iter(argv)
    for a in it:
        if a in ("-n", "--name"):
            o.name = next(it)
        elif a in ("-a", "--anchor"):
            o.asize = int(next(it))
        elif a in ("-m", "--margin"):
            o.margin = int(next(it))
        elif a in ("-d", "--max-mismatch"):
            o.maxdist = int(next(it))
        elif a == "--min-uniq-qual":
            o.min_uniq_qual = int(next(it))
        elif a == "--short-threshold":
            o.short_threshold = int(next(it))
        elif a == "--huge-threshold":
            o.huge_threshold = int(next(it))
        elif a == "--non-canonical":
            o.noncanonical = True
        elif a == "--all-hits":
            o.allhits = True
        elif a == "--strand-pref":
            o.strandpref = True
        elif a == "--half-unique":
            o.halfunique = True
        elif a == "--report-nobridges":
            o.report_nobridges = True
        elif a == "--no-linear":
            o.nolinear = True
        elif a == "--no-multi":
            o.multi_events = False
        elif a == "--test":
            o.test = True
        elif a == "--noop":
            o.noop = True
        elif a == "--stdout":
            o.stdout = next(it)
        elif a in ("-t", "--throughput"):
            pass  # progress on stderr only
        elif a == "--chunk-size":
            next(it)
        elif a == "--known-circ":
            o.known_circ = next(it)
        elif a == "--known-lin":
            o.known_lin = next(it)
        else:
            raise ValueError("unsupported option %r" % a)
    return o
