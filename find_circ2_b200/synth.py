"""
Synthetic genomes, planted circRNAs and anchor-pair alignments.

No aligner exists on the bench box, so the inputs of the hot path are generated: a random genome with
N runs and soft-masked stretches, back-splice and linear junctions flanked by GT/AG (CT/AC on the minus
strand), and reads across those junctions emitted as the records an aligner would have produced:

  * BWA-MEM style multi-segment records (`xMyS` primary + `xHyM` supplementary, AS/XS/NM tags) --
    what the shipped reference consumes (find_circ.py:375-381, 1450-1486);
  * bowtie2 style 20-nt anchor pairs named `<read>_A__<full read>` / `<read>_B` -- the v1.2 contract
    (unmapped2anchors.py:124-132, test_data/Makefile:46-53);
  * a pre-decoded struct-of-arrays batch (what ingest produces) for the bench.

The read-name grammar for the planted truth is the reference's own (find_circ.py:1148-1191):
`<name>___O:<chrom>:<pos>:<strand>;M:<n>;CS:<a>:<b>;M:<n>` etc.

Everything is seeded and deterministic.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def revcomp_bytes(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def revcomp(s: str) -> str:
    return revcomp_bytes(np.frombuffer(s.encode(), dtype=np.uint8)).tobytes().decode()


# ----------------------------------------------------------------------------------------------
# genome
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SynthGenome:
    names: List[str]
    seqs: List[np.ndarray]  # ASCII bytes (upper/lower case ACGT and N), one array per chromosome

    @property
    def sizes(self) -> List[int]:
        return [int(len(s)) for s in self.seqs]

    def chrom_index(self, name: str) -> int:
        return self.names.index(name)

    def write_fasta(self, path: str, width: int = 60) -> None:
        with open(path, "wb") as fh:
            for name, seq in zip(self.names, self.seqs):
                fh.write(b">" + name.encode() + b"\n")
                n = len(seq)
                full = (n // width) * width
                if full:
                    body = np.empty((n // width, width + 1), dtype=np.uint8)
                    body[:, :width] = seq[:full].reshape(-1, width)
                    body[:, width] = 10
                    fh.write(body.tobytes())
                if n > full:
                    fh.write(seq[full:].tobytes() + b"\n")


def make_genome(
    sizes: Sequence[int],
    seed: int = 1,
    n_frac: float = 0.005,
    n_run: Tuple[int, int] = (50, 5000),
    soft_frac: float = 0.02,
    soft_run: Tuple[int, int] = (100, 2000),
    prefix: str = "chr",
) -> SynthGenome:
    """i.i.d. uniform ACGT chromosomes; `n_frac` of the bases sit in N runs, `soft_frac` in lower-case runs."""
    rng = np.random.default_rng(seed)
    names, seqs = [], []
    for i, size in enumerate(sizes):
        seq = _ACGT[rng.integers(0, 4, size=size, dtype=np.uint8)]
        for frac, (lo, hi), what in ((soft_frac, soft_run, "soft"), (n_frac, n_run, "N")):
            if frac <= 0 or size < 4 * hi:
                continue
            mean = (lo + hi) / 2.0
            k = max(1, int(frac * size / mean))
            starts = rng.integers(0, size - hi, size=k)
            lens = rng.integers(lo, hi + 1, size=k)
            for s, ln in zip(starts, lens):
                if what == "soft":
                    seq[s : s + ln] |= 0x20  # lower-case
                else:
                    seq[s : s + ln] = ord("N")
        names.append("%s%d" % (prefix, i + 1))
        seqs.append(seq)
    return SynthGenome(names, seqs)


# ----------------------------------------------------------------------------------------------
# junctions
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Junctions:
    """start/end are BED-style 0-based half-open: the exon(s) between a back-splice occupy [start,end);
    for a linear junction [start,end) is the intron."""

    chrom: np.ndarray  # int32 index into genome.names
    start: np.ndarray  # int64
    end: np.ndarray  # int64
    minus: np.ndarray  # bool: transcript strand
    circ: np.ndarray  # bool: back-splice (True) or linear splice (False)

    def __len__(self) -> int:
        return len(self.chrom)


def plant_junctions(
    genome: SynthGenome,
    n_circ: int,
    n_lin: int,
    seed: int = 2,
    span: Tuple[int, int] = (200, 50000),
    margin: int = 400,
    with_signal: bool = True,
) -> Junctions:
    """Choose junctions and write the splice dinucleotides into the genome (upper case).

    back-splice on '+': genome[start-2:start]='AG', genome[end:end+2]='GT'  (find_circ.py:924-954)
    back-splice on '-': genome[start-2:start]='AC', genome[end:end+2]='CT'
    linear intron '+' : genome[start:start+2]='GT', genome[end-2:end]='AG'
    linear intron '-' : genome[start:start+2]='CT', genome[end-2:end]='AC'
    """
    rng = np.random.default_rng(seed)
    n = n_circ + n_lin
    sizes = np.asarray(genome.sizes, dtype=np.int64)
    p = sizes / sizes.sum()
    chrom = rng.choice(len(sizes), size=n, p=p).astype(np.int32)
    lo, hi = span
    # log-uniform spans
    sp = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)
    sp = np.minimum(sp, np.maximum(sizes[chrom] - 2 * margin - 4, 8))
    start = (margin + rng.random(n) * (sizes[chrom] - sp - 2 * margin)).astype(np.int64)
    start = np.maximum(start, 2)
    end = start + sp
    minus = rng.random(n) < 0.5
    circ = np.zeros(n, dtype=bool)
    circ[:n_circ] = True
    if with_signal:
        for i in range(n):
            g = genome.seqs[chrom[i]]
            s, e = int(start[i]), int(end[i])
            if circ[i]:
                left, right = (b"AC", b"CT") if minus[i] else (b"AG", b"GT")
                g[s - 2 : s] = np.frombuffer(left, dtype=np.uint8)
                g[e : e + 2] = np.frombuffer(right, dtype=np.uint8)
            else:
                left, right = (b"CT", b"AC") if minus[i] else (b"GT", b"AG")
                g[s : s + 2] = np.frombuffer(left, dtype=np.uint8)
                g[e - 2 : e] = np.frombuffer(right, dtype=np.uint8)
    return Junctions(chrom, start, end, minus, circ)


# ----------------------------------------------------------------------------------------------
# two-segment reads (the bulk of every config): vectorised
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class PairTable:
    """One row per read that an aligner split in two segments (= one anchor pair, find_circ.py:1058-1140).

    Everything is in genome orientation, as a SAM file stores it."""

    junc: np.ndarray  # int64 index into Junctions (-1 for decoys)
    chrom: np.ndarray  # int32
    a_pos: np.ndarray  # int64, 0-based start of the segment that comes FIRST in the read
    a_len: np.ndarray  # int32 aligned length of that segment
    b_pos: np.ndarray  # int64, 0-based start of the segment that comes SECOND in the read
    b_len: np.ndarray  # int32
    reverse: np.ndarray  # bool: SAM flag 0x10 on both segments
    reads: np.ndarray  # uint8 [n, read_len] ASCII
    as_a: np.ndarray  # int32 alignment score of A
    xs_a: np.ndarray  # int32 suboptimal score (-1: tag absent)
    as_b: np.ndarray
    xs_b: np.ndarray
    primary_is_b: np.ndarray  # bool: the second segment is the primary record
    name_id: np.ndarray  # int64 read serial number

    def __len__(self) -> int:
        return len(self.chrom)

    @property
    def read_len(self) -> int:
        return int(self.reads.shape[1])


def _gather_reads(genome: SynthGenome, chrom, left_pos, left_len, right_pos, read_len) -> np.ndarray:
    """reads[i] = genome[chrom][left_pos:left_pos+left_len] + genome[chrom][right_pos:right_pos+read_len-left_len],
    upper-cased; positions outside the chromosome read as N."""
    n = len(chrom)
    out = np.empty((n, read_len), dtype=np.uint8)
    col = np.arange(read_len, dtype=np.int64)[None, :]
    step = 1 << 17  # rows per chunk: bounds the [rows, read_len] int64 temporaries
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        ll = left_len[lo:hi].astype(np.int64)[:, None]
        gpos = np.where(col < ll, left_pos[lo:hi, None] + col, right_pos[lo:hi, None] + (col - ll))
        cc = chrom[lo:hi]
        for c in np.unique(cc):
            rows = np.nonzero(cc == c)[0]
            g = genome.seqs[c]
            p = gpos[rows]
            ok = (p >= 0) & (p < len(g))
            vals = g[np.clip(p, 0, len(g) - 1)] & 0xDF  # upper-case
            out[lo + rows] = np.where(ok, vals, ord("N")).astype(np.uint8)
    return out


def make_pairs(
    genome: SynthGenome,
    junctions: Junctions,
    n_pairs: int,
    read_len: int = 100,
    asize: int = 20,
    seed: int = 3,
    error_rate: float = 0.0,
    zipf: float = 1.0,
    frac_decoy: float = 0.10,
    frac_nonuniq: float = 0.02,
    frac_edge: float = 0.01,
    frac_inner_shift: float = 0.05,
    frac_read_n: float = 0.005,
    frac_no_xs: float = 0.05,
    min_anchor: Optional[int] = None,
) -> PairTable:
    """Reads across the planted junctions.  The junction sits `j` bases into the read, j uniform in
    [min_anchor, read_len-min_anchor] (default min_anchor = asize-2 so that breakpoints inside the
    anchor margin occur)."""
    rng = np.random.default_rng(seed)
    nj = len(junctions)
    if min_anchor is None:
        min_anchor = max(asize - 2, 1)
    # Zipf-like popularity over junctions
    w = 1.0 / np.power(np.arange(1, nj + 1, dtype=np.float64), zipf)
    rng.shuffle(w)
    w /= w.sum()
    jx = rng.choice(nj, size=n_pairs, p=w).astype(np.int64)

    chrom = junctions.chrom[jx].copy()
    jstart = junctions.start[jx]
    jend = junctions.end[jx]
    circ = junctions.circ[jx]
    sizes = np.asarray(genome.sizes, dtype=np.int64)

    j = rng.integers(min_anchor, read_len - min_anchor + 1, size=n_pairs).astype(np.int64)
    # back-splice: first part ends at jend, second starts at jstart
    # linear     : first part ends at jstart (intron start), second starts at jend (intron end)
    a_end = np.where(circ, jend, jstart)
    b_pos = np.where(circ, jstart, jend).astype(np.int64)
    a_pos = (a_end - j).astype(np.int64)
    a_len = j.astype(np.int32)
    b_len = (read_len - j).astype(np.int32)

    # decoys: same orientation, random second locus without splice signal (chance hits only)
    decoy = rng.random(n_pairs) < frac_decoy
    nd = int(decoy.sum())
    if nd:
        span = rng.integers(300, 20000, size=nd)
        back = rng.random(nd) < 0.7
        ap = a_pos[decoy]
        b_pos[decoy] = np.where(back, ap - span, ap + span + read_len)
        jx[decoy] = -1

    # pairs hugging a chromosome boundary: windows run into the N padding (find_circ.py:194-211)
    edge = rng.random(n_pairs) < frac_edge
    ne = int(edge.sum())
    if ne:
        csize = sizes[chrom[edge]]
        at_end = rng.random(ne) < 0.5
        jj = j[edge]
        off = rng.integers(0, 40, size=ne)
        # back-splice shaped: A near the chromosome end, B near its start (or both near one end)
        new_a_end = np.where(at_end, csize - off, 600 + off)
        a_pos[edge] = new_a_end - jj
        b_pos[edge] = np.where(at_end, new_a_end - 500, off)
        jx[edge] = -1

    # keep every segment inside its chromosome (windows that lie entirely outside a chromosome are
    # undefined behaviour in the reference: find_circ.py:196-209 returns strings of the wrong length)
    csz = sizes[chrom]
    a_pos = np.clip(a_pos, 0, csz - a_len)
    b_pos = np.clip(b_pos, 0, csz - b_len - 8)

    reads = _gather_reads(genome, chrom, a_pos, a_len, b_pos, read_len)

    # sequencing errors: substitutions only (simulate_reads.py:146-159 does the same)
    if error_rate > 0:
        mask = rng.random(reads.shape, dtype=np.float32) < error_rate
        shift = rng.integers(1, 4, size=int(mask.sum()), dtype=np.uint8)
        cur = reads[mask]
        code = np.searchsorted(_ACGT, cur)  # N -> 4 (stays N)
        isb = cur != ord("N")
        new = np.where(isb, _ACGT[(np.minimum(code, 3) + shift) & 3], cur)
        reads[mask] = new
    if frac_read_n > 0:
        rn = np.nonzero(rng.random(n_pairs) < frac_read_n)[0]
        reads[rn, rng.integers(0, read_len, size=len(rn))] = ord("N")

    # aligner over-extension: the inner boundary of the segments is not the true breakpoint
    shift = np.zeros(n_pairs, dtype=np.int64)
    sh = rng.random(n_pairs) < frac_inner_shift
    shift[sh] = rng.integers(-4, 5, size=int(sh.sum()))
    # keep both segments >= 1 and make short true anchors look >= asize as an aligner extension would
    short_a = a_len < asize
    shift[short_a] = asize - a_len[short_a]
    short_b = b_len < asize
    shift[short_b] = -(asize - b_len[short_b])
    shift = np.maximum(shift, -b_pos)  # segment B must not start before the chromosome does
    a_len2 = (a_len + shift).astype(np.int32)
    b_len2 = (read_len - a_len2).astype(np.int32)
    b_pos2 = b_pos + shift  # B keeps its end, its start moves with the boundary

    reverse = rng.random(n_pairs) < 0.5
    as_a = a_len2.copy()
    as_b = b_len2.copy()
    xs_a = np.maximum(as_a - rng.integers(2, 30, size=n_pairs), 0).astype(np.int32)
    xs_b = np.maximum(as_b - rng.integers(2, 30, size=n_pairs), 0).astype(np.int32)
    nonu = rng.random(n_pairs) < frac_nonuniq
    side = rng.random(n_pairs) < 0.5
    xs_a[nonu & side] = as_a[nonu & side] - rng.integers(0, 2, size=int((nonu & side).sum()))
    xs_b[nonu & ~side] = as_b[nonu & ~side] - rng.integers(0, 2, size=int((nonu & ~side).sum()))
    noxs = rng.random(n_pairs) < frac_no_xs
    xs_a[noxs] = -1
    xs_b[noxs & (rng.random(n_pairs) < 0.5)] = -1
    primary_is_b = b_len2 > a_len2

    return PairTable(
        junc=jx,
        chrom=chrom.astype(np.int32),
        a_pos=a_pos,
        a_len=a_len2,
        b_pos=b_pos2,
        b_len=b_len2,
        reverse=reverse,
        reads=reads,
        as_a=as_a.astype(np.int32),
        xs_a=xs_a,
        as_b=as_b.astype(np.int32),
        xs_b=xs_b,
        primary_is_b=primary_is_b,
        name_id=np.arange(n_pairs, dtype=np.int64),
    )


# ----------------------------------------------------------------------------------------------
# SAM text writers
# ----------------------------------------------------------------------------------------------
def sam_header(genome: SynthGenome) -> str:
    lines = ["@HD\tVN:1.3\tSO:unsorted"]
    for n, s in zip(genome.names, genome.sizes):
        lines.append("@SQ\tSN:%s\tLN:%d" % (n, s))
    lines.append("@PG\tID:synth\tPN:find_circ2_b200.synth")
    return "\n".join(lines) + "\n"


def _tags(as_, xs, nm=0) -> str:
    t = "NM:i:%d\tAS:i:%d" % (nm, as_)
    if xs >= 0:
        t += "\tXS:i:%d" % xs
    return t


def bwa_records_for_pair(
    genome: SynthGenome, t: PairTable, i: int, qname: str, qual: Optional[str] = None, base_flag: int = 0
) -> List[str]:
    """Two SAM lines (primary first) for row i: soft clips on the primary, hard clips on the supplementary."""
    R = t.read_len
    read = t.reads[i].tobytes().decode()
    q = qual if qual is not None else "*"
    al, bl = int(t.a_len[i]), int(t.b_len[i])
    fl = base_flag | (16 if t.reverse[i] else 0)
    cname = genome.names[int(t.chrom[i])]
    # segment A: first a_len bases of the read; segment B: last b_len bases (they tile the read)
    a_clip, b_clip = R - al, R - bl
    if not t.primary_is_b[i]:
        prim = (cname, int(t.a_pos[i]) + 1, "%dM%dS" % (al, a_clip), read, q, _tags(int(t.as_a[i]), int(t.xs_a[i])))
        supp = (
            cname,
            int(t.b_pos[i]) + 1,
            "%dH%dM" % (b_clip, bl),
            read[R - bl :],
            q if q == "*" else q[R - bl :],
            _tags(int(t.as_b[i]), int(t.xs_b[i])),
        )
    else:
        prim = (cname, int(t.b_pos[i]) + 1, "%dS%dM" % (b_clip, bl), read, q, _tags(int(t.as_b[i]), int(t.xs_b[i])))
        supp = (
            cname,
            int(t.a_pos[i]) + 1,
            "%dM%dH" % (al, a_clip),
            read[:al],
            q if q == "*" else q[:al],
            _tags(int(t.as_a[i]), int(t.xs_a[i])),
        )
    out = []
    for k, (c, pos, cig, seq, ql, tg) in enumerate((prim, supp)):
        f = fl | (2048 if k else 0)
        out.append("%s\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t%s\t%s\n" % (qname, f, c, pos, cig, seq, ql, tg))
    return out


def write_bwa_sam(genome: SynthGenome, t: PairTable, path: str, name_prefix: str = "r", with_qual: bool = False) -> None:
    with open(path, "w") as fh:
        fh.write(sam_header(genome))
        qual = "I" * t.read_len if with_qual else None
        for i in range(len(t)):
            fh.writelines(bwa_records_for_pair(genome, t, i, "%s%d" % (name_prefix, int(t.name_id[i])), qual))


def anchor_records_for_pair(genome: SynthGenome, t: PairTable, i: int, qname: str, asize: int) -> List[str]:
    """bowtie2-style records for the two `asize`-nt anchors of read i (v1.2 contract).

    unmapped2anchors.py:124-132 names them `<q>_A__<full read as sequenced>` and `<q>_B`; for a read
    that maps to the minus strand bowtie2 reports the reverse complement of each anchor with flag 16,
    and the anchor that comes FIRST in the read (A) lies downstream in the genome."""
    R = t.read_len
    gread = t.reads[i].tobytes().decode()  # genome orientation
    cname = genome.names[int(t.chrom[i])]
    left_pos = int(t.a_pos[i])  # genome-left anchor: first `asize` bases of the genome-oriented read
    right_pos = int(t.b_pos[i]) + int(t.b_len[i]) - asize
    left_seq, right_seq = gread[:asize], gread[R - asize :]
    left_tags = _tags(0 if t.xs_a[i] != t.as_a[i] else 0, -1 if t.xs_a[i] < 0 else int(t.xs_a[i]) - int(t.as_a[i]))
    right_tags = _tags(0 if t.xs_b[i] != t.as_b[i] else 0, -1 if t.xs_b[i] < 0 else int(t.xs_b[i]) - int(t.as_b[i]))
    if not t.reverse[i]:
        orig = gread
        recs = [
            ("%s_A__%s" % (qname, orig), 0, left_pos, left_seq, left_tags),
            ("%s_B" % qname, 0, right_pos, right_seq, right_tags),
        ]
    else:
        orig = revcomp(gread)
        # as sequenced, anchor A = first asize bases of `orig` = RC of the genome-right anchor
        recs = [
            ("%s_A__%s" % (qname, orig), 16, right_pos, right_seq, right_tags),
            ("%s_B" % qname, 16, left_pos, left_seq, left_tags),
        ]
    out = []
    for name, fl, pos, seq, tg in recs:
        # bowtie2 end-to-end scores are <= 0; keep AS:i:0 for a perfect anchor and XS <= AS
        tg = tg.replace("XS:i:", "XS:i:")
        out.append("%s\t%d\t%s\t%d\t42\t%dM\t*\t0\t0\t%s\t%s\t%s\n" % (name, fl, cname, pos + 1, asize, seq, "I" * asize, tg))
    return out


# ----------------------------------------------------------------------------------------------
# hand-assembled fragments (multi-segment / paired-end cases of record_hits, find_circ.py:1276-1439)
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Seg:
    """One aligned piece of a read, genome orientation: read[q0:q1] aligns at chrom:pos (ungapped)."""

    chrom: str
    pos: int
    q0: int
    q1: int
    reverse: bool = False
    as_: Optional[int] = None
    xs: Optional[int] = None


def records_for_read(
    qname: str, read: str, segs: Sequence[Seg], mate: int = 0, qual: Optional[str] = None, unmapped: bool = False
) -> List[str]:
    """SAM lines for one read: segs[0] is the primary (full sequence, soft clips), the others are
    supplementary (hard clips, trimmed sequence).  mate: 0 unpaired, 1/2 for paired-end flags."""
    pf = {0: 0, 1: 0x1 | 0x40, 2: 0x1 | 0x80}[mate]
    R = len(read)
    q = qual if qual is not None else "*"
    if unmapped:
        return ["%s\t%d\t*\t0\t0\t*\t*\t0\t0\t%s\t%s\n" % (qname, pf | 4, read, q)]
    out = []
    for k, s in enumerate(segs):
        clip = "H" if k else "S"
        cig = ""
        if s.q0:
            cig += "%d%s" % (s.q0, clip)
        cig += "%dM" % (s.q1 - s.q0)
        if R - s.q1:
            cig += "%d%s" % (R - s.q1, clip)
        seq = read if k == 0 else read[s.q0 : s.q1]
        ql = q if (k == 0 or q == "*") else q[s.q0 : s.q1]
        as_ = s.as_ if s.as_ is not None else (s.q1 - s.q0)
        xs = s.xs if s.xs is not None else max(as_ - 10, 0)
        f = pf | (16 if s.reverse else 0) | (2048 if k else 0)
        out.append(
            "%s\t%d\t%s\t%d\t60\t%s\t*\t0\t0\t%s\t%s\t%s\n" % (qname, f, s.chrom, s.pos + 1, cig, seq, ql, _tags(as_, xs))
        )
    return out


def genome_slice(genome: SynthGenome, chrom: str, start: int, end: int) -> str:
    g = genome.seqs[genome.chrom_index(chrom)]
    return (g[start:end] & 0xDF).tobytes().decode()
