"""
unmapped2anchors -- the step in front of the fmt-1.2 path: cut the two anchors of every unmapped read.

Host tool with the command line and output format of the reference's unmapped2anchors.py (:51-60, 91-132, 168-172):
for every unmapped record whose two ends pass the quality filter, two FASTQ records

    @<qname>_A__<full read>     first  `asize` bases / qualities
    @<qname>_B                  last   `asize` bases / qualities

go to stdout; the aligner's output for them is what find_circ (fmt 1.2, `find_circ2_b200.v12`) reads.  The full read
travels in the name of anchor A (unmapped2anchors.py:124) -- that is where RunV12 recovers it from.

Reference behaviour kept on purpose:
  * the quality filter works on `uint8(qual) - 35` WITHOUT widening (unmapped2anchors.py:97): characters below '#'
    wrap around to large values and pass;
  * `-r A|B|R` reverse (not complement) anchor A, anchor B or the whole read+qualities, `-r C` reverse-complements
    the read, `-r N` is the identity, `-r P` shuffles anchor/inner parts between reads after a burn-in of 100 reads
    (a randomised control: same procedure, but python3's generator gives a different stream than python2's);
  * `-R` (sites.reads input) fails in the reference with an AttributeError on its first line (:138) and `-F` needs a
    module that is not part of the repository (:150): `-R` is rejected here, `-F` reads plain FASTA (name -> name_<n>,
    qualities 'b', unmapped2anchors.py:152-165).
Input: BAM (own BGZF/BAM reader, find_circ2_b200.samio) as in the reference; names ending in 'sam' are read as SAM text.
"""
import random
import sys
from optparse import OptionParser


from . import samio

_COMP = str.maketrans("atcgkmryswbvhdnATCGKMRYSWBVHDN", "tagcmkyrswvbdhnTAGCMKYRSWVBDHN")

USAGE = """

  %prog <alignments.bam> > unmapped_anchors.qfa

Cut anchor sequences from both ends of the unmapped reads (optionally reversed / permuted as a control).
"""


def rev_comp(seq: str) -> str:
    for ch in seq:
        if ch not in "atcgkmryswbvhdnATCGKMRYSWBVHDN":
            raise KeyError(ch)  # the reference's complement table has no other letters (:8-42)
    return seq.translate(_COMP)[::-1]


def build_parser() -> OptionParser:
    p = OptionParser(usage=USAGE)
    p.add_option("-a", "--anchor", dest="asize", type=int, default=20, help="anchor length [20]")
    p.add_option("-q", "--minqual", dest="minqual", type=int, default=5,
                 help="smallest mean quality accepted on either anchor [5]")
    p.add_option("-r", "--rev", dest="rev", type="choice", choices=["A", "B", "R", "N", "C", "P"], default="N",
                 help="control: reverse anchor A, anchor B, the read (R), reverse-complement (C), permute parts (P), none (N)")
    p.add_option("-R", "--reads", dest="reads", action="store_true", default=False,
                 help="sites.reads input (fails in the reference; rejected)")
    p.add_option("-F", "--fasta", dest="fasta", action="store_true", default=False, help="input is a FASTA file")
    return p


class Anchors(object):
    """handle_read of unmapped2anchors.py:91-132 as an object (state: the permutation pools of -r P)"""

    N_PERM = 100

    def __init__(self, asize=20, minqual=5, rev="N", out=None, rng=None):
        self.asize, self.minqual, self.rev = asize, minqual, rev
        self.out = out if out is not None else sys.stdout
        ident, back = (lambda x: x), (lambda x: x[::-1])
        self.read_f, self.a_f, self.b_f = {"A": (ident, back, ident), "B": (ident, ident, back), "R": (back, ident, ident),
                                           "N": (ident, ident, ident), "P": (ident, ident, ident),
                                           "C": (ident, ident, ident)}[rev]
        self.pool_a, self.pool_i, self.pool_b, self.burn_in = [], [], [], []
        self.rng = rng or random

    def _pick(self, pool):
        return pool.pop(self.rng.randint(0, len(pool) - 1))

    def handle(self, qname, seq, qual, unmapped=True, replay=False):
        if not unmapped:
            return
        if seq is None or qual is None:
            raise ValueError("read %s has no sequence or no qualities" % qname)
        a = self.asize
        seq, qual = self.read_f(seq), self.read_f(qual)
        # mean of (quality byte - 35) as unsigned bytes over either anchor: the subtraction wraps below '#', like the reference's
        # numpy arithmetic (unmapped2anchors.py:100-104); plain integer sums, no array per read
        qb = qual.encode("latin-1")
        head, tail = qb[:a], qb[-a:]
        if head and tail:  # (an empty read: numpy's mean is nan there, which is not smaller than anything)
            sh = sum(head) - 35 * len(head) + (256 * sum(1 for c in head if c < 35) if min(head) < 35 else 0)
            st = sum(tail) - 35 * len(tail) + (256 * sum(1 for c in tail if c < 35) if min(tail) < 35 else 0)
            if sh < self.minqual * len(head) or st < self.minqual * len(tail):
                return
        if self.rev == "P":
            self.pool_a.append((seq[:a], qual[:a]))
            self.pool_b.append((seq[-a:], qual[-a:]))
            self.pool_i.append((seq[a:-a], qual[a:-a]))
            if not replay and len(self.burn_in) < self.N_PERM:
                self.burn_in.append((qname, seq, qual))
                return
            (sa, qa), (sb, qb), (si, qi) = self._pick(self.pool_a), self._pick(self.pool_b), self._pick(self.pool_i)
            seq, qual = sa + si + sb, qa + qi + qb
        if self.rev == "C":
            seq, qual = rev_comp(seq), qual[::-1]
        w = self.out.write
        w("@%s_A__%s\n%s\n+\n%s\n" % (qname, seq, self.a_f(seq[:a]), self.a_f(qual[:a])))
        w("@%s_B\n%s\n+\n%s\n" % (qname, self.b_f(seq[-a:]), self.b_f(qual[-a:])))

    def finish(self):
        """the burn-in reads are handled once more at the end (unmapped2anchors.py:171-172)"""
        for qname, seq, qual in self.burn_in:
            self.handle(qname, seq, qual, replay=True)


def fasta_records(fh):
    name, parts = None, []
    for line in fh:
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            if name is not None:
                yield name, "".join(parts)
            name, parts = line[1:], []
        elif line:
            parts.append(line)
    if name is not None:
        yield name, "".join(parts)


def main(argv=None, out=None):
    o, args = build_parser().parse_args(list(sys.argv[1:] if argv is None else argv))
    if o.reads:
        raise SystemExit("-R/--reads: this mode fails on its first input line in the reference (unmapped2anchors.py:138)")
    if not args:
        raise SystemExit("need one input file (BAM of unmapped reads, or FASTA with -F)")
    anchors = Anchors(o.asize, o.minqual, o.rev, out)
    if o.fasta:
        with open(args[0]) as fh:
            for n, (name, seq) in enumerate(fasta_records(fh), 1):
                anchors.handle("%s_%d" % (name.replace(" ", "_"), n), seq, "b" * len(seq))
    else:
        bam_text = None
        if not args[0].endswith("sam"):
            # BAM through the C++ reader (csrc/bam.cu: BGZF inflated on several threads, records as SAM text lines); without the
            # library (a machine that only cuts anchors) the python reader does the same
            try:
                from .ingest import BamText

                bam_text = BamText(args[0])
            except (OSError, ImportError, RuntimeError):
                bam_text = None
        if bam_text is not None:
            try:
                while True:
                    chunk = bam_text.read(16 << 20)
                    if not chunk:
                        break
                    for line in chunk.decode("latin-1").split("\n"):
                        if not line:
                            continue
                        f = line.split("\t", 11)
                        anchors.handle(f[0], None if f[9] == "*" else f[9], None if f[10] == "*" else f[10], (int(f[1]) & 4) != 0)
            finally:
                bam_text.close()
        else:
            _names, _lengths, records = samio.open_alignments(args[0])
            for r in records:
                anchors.handle(r.qname, r.seq, r.qual, r.is_unmapped)
    anchors.finish()
    return 0


if __name__ == "__main__":
    sys.exit(main())
