"""
The v1.2 face of find_circ.py (README.md:41-58, 153-164, 172-225, 283-337; test_data/Makefile:46-53): bowtie2 alignments
of 20-nt read anchors named `<read>_A__<full read>` / `<read>_B` (unmapped2anchors.py:124-132) in, splice-site BED on
stdout, spliced reads (-R) and run statistics (-s) out.

The v1.2 source is NOT part of the reference tree; what pins this module is the README's option and
column tables, the Makefile invocation and ONE golden line (test_data/cdr1as_reference.bed:2).  Everything beyond that --
tie order among equal breakpoints (upstream drew a random number), the row order (python-2 dict order), the exact text of
the reads and stats files -- is "parity unpinned" and documented as such in DESIGN.md.  The breakpoint arithmetic is the
same as the shipped script's (A_flank from A.aend - margin, B_flank up to B.pos + margin are the windows of
find_circ.py:900-902 for ungapped anchors of length asize), so the same CUDA kernels serve both faces.
"""
from __future__ import annotations

import dataclasses
import sys
import time
from collections import defaultdict
from typing import Dict, Iterable, List, Optional

import numpy as np

from ._lib import HIT_DTYPE
from .engine import Engine, decode_signal
from .samio import Alignment

_COMP = bytes.maketrans(b"ACGTNacgtnKMRYSWBVHDkmryswbvhd", b"TGCANtgcanMKYRSWVBDHmkyrswvbdh")


def rev_comp(s: str) -> str:
    return s.encode("latin-1").translate(_COMP)[::-1].decode("latin-1")


@dataclasses.dataclass
class OptionsV12:
    """README.md:283-337"""

    genome: str = ""
    name: str = "unknown"
    prefix: str = ""
    min_uniq_qual: int = 2
    asize: int = 20
    margin: int = 2
    maxdist: int = 2
    wiggle: int = 2
    noncanonical: bool = False
    allhits: bool = False
    halfunique: bool = False
    report_nobridges: bool = False
    reads: Optional[str] = None
    stats: str = "runstats.log"
    batch_pairs: int = 1 << 18
    device: int = 0


BED_HEADER_V12 = [
    "chrom", "start", "end", "name", "n_reads", "strand", "n_uniq", "uniq_bridges", "best_qual_left", "best_qual_right",
    "spliced_at_begin", "spliced_at_end", "tissues", "tiss_counts", "edits", "anchor_overlap", "breakpoints", "signal",
    "strandmatch", "category",
]  # test_data/cdr1as_reference.bed:1


class RunV12(object):
    def __init__(self, opt: OptionsV12, chrom_names, engine: Optional[Engine] = None):
        self.opt = opt
        # v1.2 ranks breakpoints by (edits, anchor overlap) only; mode 2 drops the canonical bonus of the shipped score
        self.eng = engine or Engine(opt.device, opt.asize, opt.margin, opt.maxdist, 2 if opt.noncanonical else 0, False)
        self.own_engine = engine is None
        if not self.eng.chrom_names:
            self.eng.load_genome_fasta(opt.genome)
        self.sam_chroms = list(chrom_names)
        self._tid2gid: Dict[int, int] = {}
        self.N: Dict[str, int] = defaultdict(int)
        self.eff = opt.asize - opt.margin
        self.idx_base = 0
        self.n_pairs_scanned = 0
        self.reads_by_key: Dict[tuple, List[tuple]] = defaultdict(list)
        self._reset()
        self.eng.agg_reset()

    def _reset(self):
        self.b = dict(chrom=[], a=[], b=[], l=[], fl=[], internal=[], qa=[], qb=[], rh=[], qh=[], meta=[])

    def gid(self, tid):
        g = self._tid2gid.get(tid)
        if g is None:
            g = self._tid2gid[tid] = self.eng.chrom_id(self.sam_chroms[tid])
        return g

    def add_pair(self, A: Alignment, B: Alignment):
        opt, N, eff = self.opt, self.N, self.eff
        N["total"] += 1
        if A.is_unmapped or B.is_unmapped:
            N["unmapped"] += 1
            return
        if A.tid != B.tid:
            N["other_chrom"] += 1
            return
        if A.is_reverse != B.is_reverse:
            N["other_strand"] += 1
            return
        dist = B.pos - A.pos
        if abs(dist) < opt.asize:
            N["overlapping_anchors"] += 1
            return
        rev = A.is_reverse
        if (rev and dist > 0) or (not rev and dist < 0):
            circ = True
        elif (rev and dist < 0) or (not rev and dist > 0):
            circ = False
        else:
            N["fallout"] += 1
            return
        # the full read travels in the name of anchor A (unmapped2anchors.py:124)
        read = A.qname.split("__")[1]
        readname = A.qname.split("__")[0]
        if readname.endswith("_A"):
            readname = readname[:-2]
        if rev:
            A, B = B, A
            read = rev_comp(read)
        L = len(read)
        b = self.b
        b["chrom"].append(self.gid(A.tid))
        b["a"].append(A.aend - opt.margin)
        b["b"].append(B.pos + opt.margin)
        b["l"].append(L - 2 * eff)
        b["fl"].append((1 if circ else 0) | (2 if rev else 0))
        b["internal"].append(read[eff:max(L - eff, 0)].encode("latin-1"))
        dflt = -2 * opt.asize
        b["qa"].append(A.AS - (A.XS if A.XS is not None else dflt))
        b["qb"].append(B.AS - (B.XS if B.XS is not None else dflt))
        b["rh"].append(self.eng.hash_read(read.encode("latin-1")))
        b["qh"].append(self.eng.hash_bytes(readname.encode("latin-1")))
        b["meta"].append((circ, readname, read))
        if len(b["chrom"]) >= opt.batch_pairs:
            self.flush()

    def flush(self):
        b = self.b
        n = len(b["chrom"])
        if not n:
            return
        width = (max(1, max(len(x) for x in b["internal"])) + 15) // 16 * 16
        internal = np.zeros((n, width), dtype=np.uint8)
        for i, x in enumerate(b["internal"]):
            if x:
                internal[i, :len(x)] = np.frombuffer(x, dtype=np.uint8)
        hits = np.zeros(n, dtype=HIT_DTYPE)
        clip = lambda v: np.clip(np.array(v, np.int64), -32768, 32767).astype(np.int16)  # noqa: E731
        self.eng.batch_host(np.array(b["chrom"], np.int32), np.array(b["a"], np.int32), np.array(b["b"], np.int32),
                            np.array(b["l"], np.int32), np.array(b["fl"], np.uint8), internal, np.ones(n, np.uint8),
                            clip(b["qa"]), clip(b["qb"]), np.array(b["rh"], np.uint64), np.array(b["qh"], np.uint64),
                            self.idx_base, emit=not self.opt.allhits, out=hits)
        self.n_pairs_scanned += n
        if self.opt.allhits:
            raise NotImplementedError("--allhits is not wired for the v1.2 face yet")
        nh = hits["w2"] & 0xFFFF
        for i in range(n):
            circ, readname, read = b["meta"][i]
            if nh[i] == 0:
                self.N["circ_no_bp" if circ else "splice_no_bp"] += 1
                continue
            self.N["circ_reads" if circ else "spliced_reads"] += 1
            key = (b["chrom"][i], int(hits["start"][i]), int(hits["end"][i]), "-" if int(hits["w3"][i]) & 1 else "+", 0 if circ else 1)
            self.reads_by_key[key].append((readname, read))
        self.idx_base += n
        self._reset()

    def process(self, records: Iterable[Alignment]):
        it = iter(records)
        for A in it:
            try:
                B = next(it)
            except StopIteration:
                break
            self.add_pair(A, B)
        self.flush()

    # ------------------------------------------------------------------ output
    def finalize(self):
        nj = self.eng.agg_finalize(0)
        self.junctions = self.eng.agg_fetch(nj)
        return self.junctions

    def _categories(self, r, kind) -> List[str]:
        """README.md:201-225"""
        cats = ["LINEAR" if kind else "CIRCULAR"]
        if decode_signal((int(r["sk"]) >> 16) & 0xFFF) == "GTAG":
            cats.append("CANONICAL")
        if int(r["best_q_left"]) > 0 and int(r["best_q_right"]) > 0:
            cats.append("ANCHOR_UNIQUE")
        if float(r["n_uniq_bridges"]) == 0:
            cats.append("NO_UNIQ_BRIDGES")
        if int(r["min_n_hits"]) == 1:
            cats.append("UNAMBIGUOUS_BP")
        ov, ed = int(r["min_ov"]), int(r["min_dist"])
        if ov == 0 and ed == 0:
            cats.append("PERFECT_EXT")
        elif ov <= 1 and ed <= 1:
            cats.append("GOOD_EXT")
        elif ov <= 2 and ed <= 2:
            cats.append("OK_EXT")
        return cats

    def bed_and_reads(self):
        """(BED text for stdout, reads FASTA text, stats text)"""
        opt = self.opt
        cn = self.eng.chrom_names
        junc = self.junctions
        # linear splice reads per (chrom, position, strand): the spliced_at_begin / spliced_at_end columns count reads of
        # linear junctions that use a splice site within +-wiggle of a junction's ends (merge_bed.py:123-124 sums them)
        lin_sites: Dict[tuple, int] = defaultdict(int)
        for r in junc:
            if (int(r["sk"]) >> 1) & 1:
                strand = "-" if int(r["sk"]) & 1 else "+"
                lin_sites[(int(r["chrom"]), int(r["start"]), strand)] += int(r["n_spanned"])
                lin_sites[(int(r["chrom"]), int(r["end"]), strand)] += int(r["n_spanned"])
        bed, reads = [], []
        for kind, tag in ((0, "circ"), (1, "norm")):
            bed.append("# " + "\t".join(BED_HEADER_V12) + "\n")
            n = 0
            for r in junc:
                sk = int(r["sk"])
                if ((sk >> 1) & 1) != kind:
                    continue
                ql, qr = int(r["best_q_left"]), int(r["best_q_right"])
                if opt.halfunique:
                    if ql < opt.min_uniq_qual and qr < opt.min_uniq_qual:
                        self.N["anchor_not_uniq"] += 1
                        continue
                elif ql < opt.min_uniq_qual or qr < opt.min_uniq_qual:
                    self.N["anchor_not_uniq"] += 1
                    continue
                bridges = int(round(float(r["n_uniq_bridges"])))
                if bridges == 0 and not opt.report_nobridges:
                    self.N["no_uniq_bridges"] += 1
                    continue
                n += 1
                name = "%s%s_%06d" % (opt.prefix, tag, n)
                start, end = int(r["start"]), int(r["end"])
                strand = "-" if sk & 1 else "+"
                gid = int(r["chrom"])
                at_begin = sum(lin_sites.get((gid, x, strand), 0) for x in range(start - opt.wiggle, start + opt.wiggle + 1))
                at_end = sum(lin_sites.get((gid, x, strand), 0) for x in range(end - opt.wiggle, end + opt.wiggle + 1))
                if kind == 1:
                    at_begin -= int(r["n_spanned"])
                    at_end -= int(r["n_spanned"])
                n_reads = int(r["n_spanned"])
                cols = [cn[gid], start, end, name, n_reads, strand, int(r["n_uniq"]), bridges, ql, qr, at_begin, at_end,
                        opt.name, n_reads, int(r["min_dist"]), int(r["min_ov"]), int(r["min_n_hits"]),
                        decode_signal((sk >> 16) & 0xFFF), "NA", ",".join(sorted(self._categories(r, kind)))]
                bed.append("\t".join(str(c) for c in cols) + "\n")
                for readname, read in self.reads_by_key.get((gid, start, end, strand, kind), []):
                    reads.append(">%s %s\n%s\n" % (readname, name, read))
        stats = "".join("%s\t%d\n" % (k, self.N[k]) for k in sorted(self.N))
        return "".join(bed), "".join(reads), stats

    def close(self):
        if self.own_engine:
            self.eng.close()
