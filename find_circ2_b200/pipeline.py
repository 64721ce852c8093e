"""
Host side of the drop-in: the reference's main loop (find_circ.py:1490-1610) re-organised around batched GPU calls.

  alignments -> fragments (find_circ.py:1450-1486) -> mates / adjacent segment pairs (:976-1140)
             -> batches of anchor pairs (struct of arrays) -> Engine.batch_host(): breakpoint scan on the GPU
             -> per-fragment evidence logic on the returned hits (record_hits, :1276-1439; control flow only)
             -> junction records appended to the GPU aggregator -> sort/reduce on the GPU
             -> BED / reads / multi-event / counter writers (:695-730, 733-763, 1442-1447, 1605-1607)

Nothing here computes a breakpoint or a junction count: with no CUDA device the Engine constructor raises.
Output row order: the reference iterates a python2 dict (hash order); here rows come in discovery order, which is the
order python3 would give -- comparisons are made after a canonical sort (BASELINE.json north_star).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import gzip
import logging
import os
import sys
import time
from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import HIT_DTYPE, JREC_DTYPE
from .engine import Engine, decode_signal
from .samio import Alignment

VERSION = "1.99-b200"


@dataclasses.dataclass
class Options:
    """find_circ.py:383-413"""

    genome: str = ""
    output: str = "find_circ_run"
    name: str = "unknown"
    min_uniq_qual: int = 2
    asize: int = 15
    margin: int = 2
    maxdist: int = 2
    short_threshold: int = 100
    huge_threshold: int = 100000
    noncanonical: bool = False
    allhits: bool = False
    strandpref: bool = False
    halfunique: bool = False
    report_nobridges: bool = False
    nolinear: bool = False
    multi_events: bool = True
    throughput: bool = False
    chunksize: int = 100000
    noop: bool = False
    silent: bool = False
    stdout: Optional[str] = None
    batch_pairs: int = 1 << 18  # anchor pairs per GPU batch (ours)
    device: int = 0
    native: bool = True  # native (C++) SAM ingest for SAM text files
    ingest_threads: int = 0  # parser threads of the native ingest (0: one per core, at most 16)
    ingest_piece_bytes: int = 4 << 20  # chunks of SAM text larger than this are cut into one piece per thread
    test: bool = False    # test_results.tsv: every fragment against the truth in its name (find_circ.py:411, 1148-1273)
    known_circ: str = ""  # BED6 files of known junctions (find_circ.py:387-388)
    known_lin: str = ""


def py2_str(x) -> str:
    """python2 str(): 12 significant digits for floats"""
    if isinstance(x, (bool, np.bool_)):
        return "True" if x else "False"
    if isinstance(x, (float, np.floating)):
        s = "%.12g" % float(x)
        if s in ("inf", "-inf", "nan"):
            return s
        if "." not in s and "e" not in s:
            s += ".0"
        return s
    return str(x)


class Mate(object):
    """MateSegments (find_circ.py:976-1056)"""

    __slots__ = ("primary", "proper", "other_chrom", "other_strand")

    def __init__(self, primary: Alignment):
        self.primary = primary
        self.proper = [primary]
        self.other_chrom: List[Alignment] = []
        self.other_strand: List[Alignment] = []

    def add(self, rec: Alignment):
        if rec.tid != self.primary.tid:
            self.other_chrom.append(rec)
        elif rec.is_reverse != self.primary.is_reverse:
            self.other_strand.append(rec)
        else:
            self.proper.append(rec)


class Span(object):
    """one anchor pair = JunctionSpan (find_circ.py:821-852); `row` is its index in the current GPU batch"""

    __slots__ = ("A", "B", "primary", "q_start", "q_end", "den", "uniq", "backsplice", "row")

    def __init__(self, A, B, primary, q_start, q_end, den):
        self.A = A
        self.B = B
        self.primary = primary
        self.q_start = q_start
        self.q_end = q_end
        self.den = den
        self.uniq = min(A.uniqueness(), B.uniqueness())
        self.backsplice = (B.pos - A.aend) < 0
        self.row = -1


class Fragment(object):
    __slots__ = ("name", "mates", "circ", "lin", "unspliced", "broken", "conditional", "seq")

    def __init__(self, name):
        self.name = name
        self.mates: List[Mate] = []
        self.circ: List[Span] = []
        self.lin: List[Span] = []
        self.unspliced: List[Alignment] = []
        self.broken: List[Alignment] = []
        self.conditional = False


def iter_fragments(records: Iterable[Alignment], N) -> Iterable[Tuple[Optional[Mate], Mate]]:
    """find_circ.py:1450-1486"""
    it = iter(records)
    try:
        first = next(it)
    except StopIteration:
        return
    N["total_mates"] += 1
    cur, other = Mate(first), None
    for rec in it:
        if rec.flag & 0x4:
            N["unmapped_reads"] += 1
            continue
        p = cur.primary
        if rec.qname == p.qname:
            if (rec.flag & 0x40) == (p.flag & 0x40):
                cur.add(rec)
            else:
                N["total_mates"] += 1
                other, cur = cur, Mate(rec)
        else:
            yield other, cur
            N["total_mates"] += 1
            other, cur = None, Mate(rec)
    yield other, cur


def mate_spans(mate: Mate, asize: int, N) -> List[Span]:
    """adjacent_segment_pairs (find_circ.py:1058-1140)"""
    segs = mate.proper
    den = len(segs) - 1
    starts = [s.clip_start() for s in segs]
    ends = [st + s.query_length() for st, s in zip(starts, segs)]
    order = sorted(range(len(segs)), key=lambda k: starts[k])
    out = []
    for ia, ib in zip(order, order[1:]):
        if ends[ia] - starts[ia] < asize or ends[ib] - starts[ib] < asize:
            N["seg_too_short_skip"] += 1
            continue
        out.append(Span(segs[ia], segs[ib], mate.primary, min(starts[ia], starts[ib]), max(ends[ia], ends[ib]), den))
    return out


def truth_from_name(text: str):
    """the fragment structure that a simulated read carries behind '___' in its name (find_circ.py:1147-1191): mates are
    separated by '|', steps by ';'.  O:chrom:start:strand sets the origin, M:n walks n bases, LS:a:b / CS:a:b are a linear /
    circular splice with both ends relative to the origin.  Origin and position carry over from mate to mate, as upstream."""
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
            elif f[0] in ("LS", "CS"):
                left, right = int(f[1]) + start, int(f[2]) + start
                (lin if f[0] == "LS" else circ).add((chrom, left, right, strand))
                spliced, end = True, (right if f[0] == "LS" else left)
        if not spliced and chrom:
            unspliced.add((chrom, start, end, "*"))
    return lin, circ, unspliced


def test_result_row(frag: str, lin_coords, circ_coords, unspliced_coords, broken_coords) -> str:
    """one line of test_results.tsv (find_circ.py:1194-1273)"""
    if "___" not in frag:
        return "\t".join([frag, "N/A", "N/A", "N/A", "N/A"])
    lin_ref, circ_ref, un_ref = truth_from_name(frag.split("___")[-1])

    def verdict(ref, got, what, ok):
        flags = []
        if ref - got:
            flags.append("MISSED_%s:%s" % (what, ",".join(str(c) for c in sorted(ref - got))))
        if got - ref:
            flags.append("SPURIOUS_%s:%s" % (what, ",".join(str(c) for c in sorted(got - ref))))
        return ";".join(sorted(flags)) if flags else (ok if ref else "N/A")

    broken = "BROKEN_SEGMENTS:" + ";".join(str(b) for b in sorted(broken_coords)) if broken_coords else "N/A"
    return "\t".join([frag, verdict(lin_ref, set(lin_coords), "LINEAR_JUNCTIONS", "LIN_OK"),
                      verdict(circ_ref, set(circ_coords), "CIRCULAR_JUNCTIONS", "CIRC_OK"),
                      verdict(un_ref, set(unspliced_coords), "UNSPLICED", "UNSPLICED_OK"), broken])


def _unique_rows(key: np.ndarray):
    """np.unique(key, axis=0, return_inverse=True) for an [n, w] int64 matrix (junction identities, combinations of them)
    without the lexicographic sort: distinct rows in order of first appearance and the number of every row's value
    (fc_unique_rows: one exact hash table in C++)"""
    if len(key) == 0:
        return key, np.zeros(0, dtype=np.int64)
    from . import _lib

    key = np.ascontiguousarray(key, dtype=np.int64)
    n, w = key.shape
    first = np.empty(n, dtype=np.int64)
    inverse = np.empty(n, dtype=np.int32)
    nu = _lib.load().fc_unique_rows(key.ctypes.data, n, w, first.ctypes.data, inverse.ctypes.data)
    if nu < 0:
        raise RuntimeError("fc_unique_rows failed (%d)" % nu)
    return key[first[:nu]], inverse


def _piece_cuts(buf, pieces: int, start: int = 0, end: Optional[int] = None):
    """offsets that cut the SAM text buf[start:end] into `pieces` parts on fragment boundaries (a line whose read name
    differs from the line before it, find_circ.py:1450-1486): the parts can be parsed independently"""
    n = len(buf) if end is None else end
    cuts = [start]
    for k in range(1, pieces):
        pos = buf.find(b"\n", max(start + (n - start) * k // pieces, cuts[-1]), n)
        if pos < 0:
            break
        ls = pos + 1  # start of a line; skip the rest of its fragment
        if ls >= n:
            break
        name = buf[ls:buf.find(b"\t", ls, n)]
        while True:
            nl = buf.find(b"\n", ls, n)
            if nl < 0:
                ls = n
                break
            ls = nl + 1
            if ls >= n or buf[ls:buf.find(b"\t", ls, n)] != name:
                break
        if ls >= n:
            break
        if ls > cuts[-1]:
            cuts.append(ls)
    cuts.append(n)
    return cuts


# per-fragment evidence flags of record_hits (find_circ.py:1331-1437) as bits, in the order sorted() gives their names
FLAG_NAMES = sorted(["BROKEN_SEGMENTS", "SUPPORT_CLOSURE", "SUPPORT_INSIDE_MATE", "SUPPORT_INSIDE_SPLICE_JUNCTION",
                     "WARN_MULTI_BACKSPLICE", "WARN_OTHER_CHROM_MATE", "WARN_OUTSIDE_MATE", "WARN_OUTSIDE_SPLICE_JUNCTION",
                     "WARN_UNRESOLVED_EXTRA_BACKSPLICE", "WARN_UNRESOLVED_LINSPLICE"])
FLAG_BIT = {n: 1 << k for k, n in enumerate(FLAG_NAMES)}
FLAGS_NOT_WARN = sum(b for n, b in FLAG_BIT.items() if not n.startswith("WARN"))
_FLAG_TEXT: Dict[int, str] = {}


def flag_names(mask: int) -> Tuple[str, ...]:
    return tuple(n for k, n in enumerate(FLAG_NAMES) if (mask >> k) & 1)


def flag_text(mask: int) -> str:
    t = _FLAG_TEXT.get(mask)
    if t is None:
        t = _FLAG_TEXT[mask] = ",".join(flag_names(mask))
    return t


class JunctionInfo(object):
    """what the fragments' evidence flags add up to for one junction (the flags / read_flags dicts of find_circ.py:541-547,
    1433-1437): occurrences per flag, and over the distinct fragment names the number without BROKEN_SEGMENTS and the
    number of (name, non-WARN flag) combinations"""

    __slots__ = ("flags", "n_names", "unbroken", "unwarned")

    def __init__(self):
        self.flags: Dict[str, int] = {}
        self.n_names = self.unbroken = self.unwarned = 0


BED_HEADER = [
    "chrom", "start", "end", "name", "n_frags", "strand", "n_weight", "n_spanned", "n_uniq", "uniq_bridges",
    "best_qual_left", "best_qual_right", "tissues", "tiss_counts", "edits", "anchor_overlap", "breakpoints",
    "signal", "strandmatch", "category", "flags", "flag_counts",
]
MULTI_HEADER = ["chrom", "start", "end", "name", "score", "strand", "fragment_name", "lin_cons", "lin_incons",
                "unspliced_cons", "unspliced_incons"]


class Run(object):
    def __init__(self, opt: Options, chrom_names: Sequence[str], engine: Optional[Engine] = None):
        self.opt = opt
        self.N: Dict[str, float] = defaultdict(float)
        self.eng = engine or Engine(opt.device, opt.asize, opt.margin, opt.maxdist, opt.noncanonical, opt.strandpref)
        self.own_engine = engine is None
        if not self.eng.chrom_names:
            self.eng.load_genome_fasta(opt.genome)
        self.sam_chroms = list(chrom_names)
        # SAM tid -> genome chromosome id; unknown names raise KeyError when first used, like find_circ.py:193
        self._tid2gid: Dict[int, int] = {}
        self.eff = opt.asize - opt.margin
        self.frags: List[Fragment] = []
        self._reset_batch()
        self.idx_base = 0
        self.n_fragments = 0
        self.n_pairs_scanned = 0
        # evidence flags: one event (junction key, fragment-name hash, flag mask) per fragment and junction; summed up per
        # junction when the tables are written (junction_info)
        self._ev_py: List[tuple] = []      # python path: (chrom id, start, end, minus, kind, name hash, mask)
        self.ev_native: List[tuple] = []   # native path: (keys [n, 5] int64, name hashes uint64, masks uint32) per batch
        self._info_cache = None
        self.reads_out: List[tuple] = []   # (fragment ordinal, qname, seq, qual, [keys], flags)
        self.explicit_idx = False          # native ingest: rows carry their own stream position (fragment ordinal * 64 + k)
        self.cur_seq = 0
        self.cur_k = 0
        self.multi_out: List[tuple] = []   # (fragment ordinal, name, circ key, lin cons, lin incons, mate cons, mate incons)
        self.native_multi: List[dict] = []  # the same of the native ingest, arrays per batch
        self.native_reads: List[dict] = []  # spliced reads of the native ingest, one compact batch record each
        self.test_out: List[tuple] = []    # (fragment ordinal, row of test_results.tsv)
        self.t_scan = 0.0
        self.known: Dict[tuple, str] = {}
        for kind, path in ((0, opt.known_circ), (1, opt.known_lin)):
            if path:
                self._load_known(path, kind)
        self.eng.agg_reset()

    def _load_known(self, path: str, kind: int):
        """SpliceSiteStorage.load_known_sites (find_circ.py:665-679): a known site keeps its BED name, does not count
        towards the novel numbering and carries a placeholder splice (edits 10, overlap 10, one breakpoint) that takes
        part in the minima of the edits / anchor_overlap / breakpoints columns"""
        gid = {n: k for k, n in enumerate(self.eng.chrom_names)}
        for line in open(path):
            if line.startswith("#"):
                continue
            chrom, start, end, name, _score, strand = line.rstrip().split("\t")[:6]
            if chrom in gid:  # a site on a chromosome the genome does not have can never be observed
                self.known[(gid[chrom], int(start), int(end), strand, kind)] = name

    # ------------------------------------------------------------------ batch assembly
    def _reset_batch(self):
        self.b_chrom, self.b_a, self.b_b, self.b_l, self.b_fl = [], [], [], [], []
        self.b_int, self.b_den, self.b_qa, self.b_qb, self.b_rh, self.b_qh = [], [], [], [], [], []
        self.b_idx = []
        self.frags = []
        self.b_conditional = False

    def gid(self, tid: int) -> int:
        g = self._tid2gid.get(tid)
        if g is None:
            g = self.eng.chrom_id(self.sam_chroms[tid])
            self._tid2gid[tid] = g
        return g

    def _queue(self, sp: Span):
        eff = self.eff
        A, B = sp.A, sp.B
        part = sp.primary.seq[sp.q_start:sp.q_end]
        L = len(part)
        sp.row = len(self.b_chrom)
        self.b_chrom.append(self.gid(A.tid))
        self.b_a.append(A.pos + eff)
        self.b_b.append(B.aend - eff)
        self.b_l.append(L - 2 * eff)
        self.b_fl.append((1 if sp.backsplice else 0) | (2 if sp.primary.is_reverse else 0))
        self.b_int.append(part[eff:max(L - eff, 0)].encode("latin-1"))
        if sp.den > 255:
            raise ValueError("read %s has more than 256 segments" % sp.primary.qname)
        self.b_den.append(sp.den)
        # Hit.add: AS - XS with XS defaulting to 0 (find_circ.py:556-559)
        self.b_qa.append(A.AS - (A.XS or 0))
        self.b_qb.append(B.AS - (B.XS or 0))
        if self.explicit_idx:
            self.b_idx.append(self.cur_seq * 64 + min(self.cur_k, 63))
            self.cur_k += 1
        self.b_rh.append(self.eng.hash_read(sp.primary.seq.encode("latin-1")))
        self.b_qh.append(self.eng.hash_bytes(sp.primary.qname.encode("latin-1")))

    def add_fragment(self, mate1: Optional[Mate], mate2: Mate):
        """process_mate x2 + the head of record_hits (find_circ.py:1492-1526, 1560-1574)"""
        opt, N = self.opt, self.N
        self.n_fragments += 1
        if not self.explicit_idx:
            self.cur_seq = self.n_fragments
        self.cur_k = 0
        fr = Fragment(mate2.primary.qname)
        fr.seq = self.cur_seq
        for mate in (mate1, mate2):
            if mate is None:
                continue
            fr.mates.append(mate)
            if len(mate.proper) < 2:
                N["unspliced_mates"] += 1
                fr.unspliced.append(mate.primary)
                continue
            L = len(mate.primary.seq)
            lo, hi = L, 0
            for sp in mate_spans(mate, opt.asize, N):
                (fr.circ if sp.backsplice else fr.lin).append(sp)
                lo, hi = min(lo, sp.q_start), max(hi, sp.q_end)
            if hi < L - opt.asize or lo > opt.asize:
                fr.broken.extend(mate.other_chrom)
                fr.broken.extend(mate.other_strand)
        if not fr.circ and opt.nolinear:
            return
        if not (fr.circ or fr.lin):
            return
        for sp in fr.circ + fr.lin:
            if sp.uniq >= opt.min_uniq_qual:
                self._queue(sp)
        # the linear spans only count when the back-splices resolve to <= 1 junction (find_circ.py:1319-1329)
        n_circ_q = sum(1 for sp in fr.circ if sp.row >= 0)
        fr.conditional = bool(fr.lin) and (n_circ_q >= 2 or opt.nolinear or (opt.allhits and n_circ_q >= 1))
        self.b_conditional = self.b_conditional or fr.conditional
        self.frags.append(fr)
        if len(self.b_chrom) >= opt.batch_pairs:
            self.flush()

    # ------------------------------------------------------------------ GPU batch + evidence logic
    def flush(self):
        n = len(self.b_chrom)
        if not self.frags:
            return
        opt = self.opt
        two_step = self.b_conditional or opt.allhits
        hits = np.zeros(n, dtype=HIT_DTYPE)
        if n:
            width = max(1, max(len(b) for b in self.b_int))
            width = (width + 15) // 16 * 16
            internal = np.zeros((n, width), dtype=np.uint8)
            for i, b in enumerate(self.b_int):
                if b:
                    internal[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
            t0 = time.perf_counter()
            self.eng.batch_host(
                np.array(self.b_chrom, np.int32), np.array(self.b_a, np.int32), np.array(self.b_b, np.int32),
                np.array(self.b_l, np.int32), np.array(self.b_fl, np.uint8), internal,
                np.array(self.b_den, np.uint8), np.clip(np.array(self.b_qa, np.int64), -32768, 32767).astype(np.int16),
                np.clip(np.array(self.b_qb, np.int64), -32768, 32767).astype(np.int16),
                np.array(self.b_rh, np.uint64), np.array(self.b_qh, np.uint64), self.idx_base,
                emit=not two_step, out=hits, idx=np.array(self.b_idx, np.uint64) if self.explicit_idx else None)
            self.t_scan += time.perf_counter() - t0
            self.n_pairs_scanned += n
        ties_off = ties = None
        if opt.allhits and n:
            ties_off, ties = self.eng.batch_ties((hits["w2"] & 0xFFFF).astype(np.int64))
        mask = np.ones(n, dtype=np.uint8) if two_step else None
        host_recs: List[tuple] = []
        for fr in self.frags:
            self._record_hits(fr, hits, mask, ties_off, ties, host_recs)
        if two_step and n:
            if opt.allhits:
                self._append_host_records(host_recs)
            else:
                self.eng.batch_emit(mask, self.idx_base)
        self.idx_base += max(n, 1) * (64 if opt.allhits else 1)
        self._reset_batch()

    def _append_host_records(self, recs: List[tuple]):
        """--all-hits: the records (one per tie) are assembled on the host from the GPU's tie list"""
        if not recs:
            return
        arr = np.zeros(len(recs), dtype=JREC_DTYPE)
        for k, (key, row, sub, dist, ov, nh, sig) in enumerate(recs):
            gid, start, end, strand, kind = key
            rh = self.b_rh[row]
            back = self.b_fl[row] & 1
            qa, qb = self.b_qa[row], self.b_qb[row]
            arr[k] = (gid, start, end, (1 if strand == "-" else 0) | (kind << 1) | ((rh & 1) << 2) | (self.b_den[row] << 8) | (sig << 16),
                      self.idx_base + row * 64 + min(sub, 63), rh, self.b_qh[row],
                      max(-32768, min(32767, qb if back else qa)), max(-32768, min(32767, qa if back else qb)), nh, dist, ov)
        self.eng.agg_append_host(arr)

    @staticmethod
    def _hit_fields(h):
        w2, w3 = int(h["w2"]), int(h["w3"])
        return (int(h["start"]), int(h["end"]), "-" if (w3 & 1) else "+", (w2 >> 16) & 0xFF, w2 >> 24, w2 & 0xFFFF, (w3 >> 1) & 0xFFF)

    def _splices(self, sp: Span, hits, ties_off, ties):
        """the list find_breakpoints() returns, truncated to the first tie unless --all-hits (find_circ.py:1312-1317)"""
        h = hits[sp.row]
        nh = int(h["w2"]) & 0xFFFF
        if nh == 0:
            return []
        if self.opt.allhits and nh > 1:
            return [self._hit_fields(ties[k]) for k in range(int(ties_off[sp.row]), int(ties_off[sp.row + 1]))]
        return [self._hit_fields(h)]

    def _event(self, key, name: str, mask: int):
        if mask:
            gid, start, end, strand, kind = key
            self._ev_py.append((gid, start, end, 1 if strand == "-" else 0, kind, self.eng.hash_bytes(name.encode("latin-1")), mask))

    def events(self):
        """all evidence events so far as arrays: keys [n, 5] (chrom id, start, end, minus, kind), name hashes, flag masks"""
        ks = [e[0] for e in self.ev_native]
        hs = [e[1] for e in self.ev_native]
        ms = [e[2] for e in self.ev_native]
        if self._ev_py:
            ks.append(np.array([e[:5] for e in self._ev_py], dtype=np.int64).reshape(-1, 5))
            hs.append(np.array([e[5] for e in self._ev_py], dtype=np.uint64))
            ms.append(np.array([e[6] for e in self._ev_py], dtype=np.uint32))
        if not ks:
            return np.zeros((0, 5), np.int64), np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        return np.concatenate(ks), np.concatenate(hs), np.concatenate(ms)

    def junction_info(self) -> Dict[tuple, JunctionInfo]:
        """the events summed up per junction.  A fragment name stands for itself through its 64-bit hash, like in the
        fragment count of the aggregation (collision bound)."""
        n_ev = len(self._ev_py) + sum(len(e[2]) for e in self.ev_native)
        if self._info_cache is not None and self._info_cache[0] == n_ev:
            return self._info_cache[1]
        K, H, M = self.events()
        out: Dict[tuple, JunctionInfo] = {}
        if len(M):
            # group by junction, then by name: the junction of a row as one number first (two sort keys instead of six)
            _, kid = _unique_rows(K)
            kid = kid.reshape(-1)
            order = np.lexsort((H, kid))
            K, H, M, kid = K[order], H[order], M[order], kid[order]
            new_key = np.ones(len(M), dtype=bool)
            new_key[1:] = kid[1:] != kid[:-1]
            new_name = new_key.copy()
            new_name[1:] |= H[1:] != H[:-1]
            key_start = np.nonzero(new_key)[0]
            name_start = np.nonzero(new_name)[0]
            per_flag = [np.add.reduceat(((M >> k) & 1).astype(np.int64), key_start) for k in range(len(FLAG_NAMES))]
            name_mask = np.bitwise_or.reduceat(M, name_start)
            name_key = np.cumsum(new_key)[name_start] - 1  # the junction every distinct name belongs to
            n_names = np.bincount(name_key, minlength=len(key_start))
            unbroken = np.bincount(name_key, weights=(name_mask & FLAG_BIT["BROKEN_SEGMENTS"]) == 0, minlength=len(key_start))
            good = name_mask & np.uint32(FLAGS_NOT_WARN)
            pop = sum(((good >> k) & 1).astype(np.int64) for k in range(len(FLAG_NAMES)))
            unwarned = np.bincount(name_key, weights=pop, minlength=len(key_start))
            for j, k0 in enumerate(key_start.tolist()):
                gid, start, end, minus, kind = K[k0].tolist()
                inf = out[(gid, start, end, "-" if minus else "+", kind)] = JunctionInfo()
                inf.flags = {FLAG_NAMES[k]: int(per_flag[k][j]) for k in range(len(FLAG_NAMES)) if per_flag[k][j]}
                inf.n_names, inf.unbroken, inf.unwarned = int(n_names[j]), int(unbroken[j]), int(unwarned[j])
        self._info_cache = (n_ev, out)
        return out

    def _record_hits(self, fr: Fragment, hits, mask, ties_off, ties, host_recs):
        """record_hits (find_circ.py:1276-1439) on the GPU's answers; junction identity = (chrom id, start, end, strand, kind)"""
        opt, N = self.opt, self.N
        warns = set()
        junctions: List[tuple] = []

        def note(key):
            if key not in junctions:
                junctions.append(key)

        circ_coords = set()
        circ_key = None
        for sp in fr.circ:
            if sp.row < 0:
                N["circ_junc_not_unique"] += 1
                continue
            spl = self._splices(sp, hits, ties_off, ties)
            if not spl:
                N["circ_no_bp"] += 1
                warns.add("WARN_UNRESOLVED_EXTRA_BACKSPLICE")
                continue
            N["circ_spliced"] += 1
            gid = self.b_chrom[sp.row]
            for sub, (start, end, strand, dist, ov, nh, sig) in enumerate(spl):
                circ_key = (gid, start, end, strand, 0)
                if opt.allhits:
                    host_recs.append((circ_key, sp.row, sub, dist, ov, nh, sig))
                circ_coords.add(circ_key)
                note(circ_key)
                if not opt.allhits:
                    break

        def skip_linear():
            if mask is not None:
                for sp in fr.lin:
                    if sp.row >= 0:
                        mask[sp.row] = 0

        if len(circ_coords) > 1:
            for key in circ_coords:
                warns.add("WARN_MULTI_BACKSPLICE")
                self._event(key, fr.name, FLAG_BIT["WARN_MULTI_BACKSPLICE"])
                note(key)
            skip_linear()
            return self._finish_fragment(fr, junctions, warns)
        if not circ_coords and opt.nolinear:
            skip_linear()
            return self._finish_fragment(fr, junctions, warns)

        if circ_coords:
            _, circ_start, circ_end, _, _ = circ_key
            first_tid = fr.circ[0].primary.tid
            if len(fr.circ) > 1:
                warns.add("SUPPORT_CLOSURE")

        lin_cons, lin_incons = set(), set()
        lin_coords = set()
        for sp in fr.lin:
            if sp.row < 0:
                N["lin_junc_not_unique"] += 1
                continue
            spl = self._splices(sp, hits, ties_off, ties)
            if not spl:
                N["lin_no_bp"] += 1
                warns.add("WARN_UNRESOLVED_LINSPLICE")
                continue
            N["lin_spliced"] += 1
            gid = self.b_chrom[sp.row]
            cname = self.eng.chrom_names[gid]
            for sub, (start, end, strand, dist, ov, nh, sig) in enumerate(spl):
                key = (gid, start, end, strand, 1)
                if opt.allhits:
                    host_recs.append((key, sp.row, sub, dist, ov, nh, sig))
                note(key)
                lin_coords.add((cname, start, end, strand))
                if circ_coords:
                    coord = (cname, start, end, strand)
                    if start <= circ_start or end >= circ_end:
                        warns.add("WARN_OUTSIDE_SPLICE_JUNCTION")
                        lin_incons.add(coord)
                    else:
                        lin_cons.add(coord)
                        warns.add("SUPPORT_INSIDE_SPLICE_JUNCTION")
                if not opt.allhits:
                    break

        if opt.test:  # find_circ.py:1380-1394
            cn = self.eng.chrom_names
            star = lambda rec: (self.sam_chroms[rec.tid], rec.pos, rec.aend, "*")  # noqa: E731
            self.test_out.append((fr.seq, test_result_row(fr.name, lin_coords, {(cn[g], s0, e0, st) for g, s0, e0, st, _ in circ_coords},
                                                          {star(r) for r in fr.unspliced}, {star(r) for r in fr.broken})))

        if circ_coords:
            un_cons, un_incons = set(), set()
            for rec in fr.unspliced:
                coord = (self.sam_chroms[rec.tid], rec.pos, rec.aend, "*")
                if first_tid != rec.tid:
                    warns.add("WARN_OTHER_CHROM_MATE")
                    un_incons.add(coord)
                elif rec.pos + opt.asize <= circ_start or rec.aend - opt.asize >= circ_end:
                    warns.add("WARN_OUTSIDE_MATE")
                    un_incons.add(coord)
                else:
                    warns.add("SUPPORT_INSIDE_MATE")
                    un_cons.add(coord)
            if fr.broken:
                warns.add("BROKEN_SEGMENTS")
            if (un_cons or un_incons or lin_cons or lin_incons) and opt.multi_events:
                self.multi_out.append((fr.seq, fr.name, circ_key, lin_cons, lin_incons, un_cons, un_incons))
            self._event(circ_key, fr.name, sum(FLAG_BIT[w] for w in warns))
        return self._finish_fragment(fr, junctions, warns)

    def _finish_fragment(self, fr: Fragment, junctions, warns):
        if junctions:
            fl = tuple(sorted(warns))
            for mate in fr.mates:
                p = mate.primary
                self.reads_out.append((fr.seq, p.qname, p.seq, p.qual, tuple(junctions), fl))

    # ------------------------------------------------------------------ driver
    def process(self, records: Iterable[Alignment]):
        for mate1, mate2 in iter_fragments(records, self.N):
            if self.opt.noop:
                self.n_fragments += 1
                continue
            self.add_fragment(mate1, mate2)
        self.flush()

    def process_native(self, fh, chunk_bytes: int = 32 << 20, first_fragment: int = 0, at_stream_start: bool = True):
        """SAM text (binary file object, header included) through the native ingest (csrc/ingest.cu); fragments it does
        not handle go through add_fragment().  Same results as process(), an order of magnitude less host time.
        first_fragment / at_stream_start: this stream is a part of a larger one (one rank of a multi-GPU run)."""
        from .ingest import COUNTER_NAMES, NativeIngest
        from .samio import _parse_sam_line

        opt, N = self.opt, self.N
        if opt.allhits or opt.noop or opt.test:
            raise ValueError("process_native does not cover --all-hits / --noop / --test")
        tid2gid = [self.eng._chrom_ids.get(n, -1) for n in self.sam_chroms]
        name2tid = {n: i for i, n in enumerate(self.sam_chroms)}
        threads = int(getattr(opt, "ingest_threads", 0) or 0) or max(1, min(16, os.cpu_count() or 1))
        cap = max(1024, opt.batch_pairs)
        # the handles of the parser threads see one piece of a chunk each: arrays for the rows such a piece can hold (a row
        # takes more than 64 bytes of SAM text; a piece that holds more just takes another call)
        piece_cap = max(1024, min(cap, chunk_bytes // max(threads, 1) // 64 + 1024))
        mk = lambda first, at_start, c=cap: NativeIngest(opt.asize, opt.margin, opt.min_uniq_qual, opt.nolinear, self.sam_chroms, tid2gid,  # noqa: E731
                                                         cap=c, first_fragment=first, at_stream_start=at_start)
        ings = [mk(first_fragment, at_stream_start)]
        self.explicit_idx = True
        eof = False
        pool = None
        pending = []  # evidence of the batches so far, in stream order (results or futures)
        HEAD = 1 << 20  # room in front of every chunk for the tail the chunk before it left over (an unfinished fragment)

        def read_chunk():
            """the next chunk_bytes of the stream behind HEAD free bytes -> (bytearray, bytes read)"""
            if hasattr(fh, "readinto"):
                ba = bytearray(HEAD + chunk_bytes)
                got = fh.readinto(memoryview(ba)[HEAD:])
                return ba, int(got or 0)
            data = fh.read(chunk_bytes)  # (BAM text, a prefixed stdin, a rank's byte range: objects with read() only)
            ba = bytearray(HEAD + len(data))
            ba[HEAD:] = data
            return ba, len(data)

        def consume(buf, off, a, n, m, max_l, n_frag, counters, complex_ranges, plane_stride):
            """what one fc_ingest_parse call produced: rows and fragment records -> GPU + evidence rules, fragments it
            left -> python"""
            self.n_fragments += n_frag
            for k, name in enumerate(COUNTER_NAMES):
                if counters[k]:
                    N[name] += counters[k]
            if n:
                # (arrays that belong to a parser handle are overwritten by its next call: their evidence is taken right away)
                later = pool if a is not ings[0].a else None
                pending.append(self._native_batch(buf, off, a, n, m, max_l, ings[0].n_words, plane_stride, ings[0].lib, later))
                while len(pending) > 64:  # (bounded: a future holds its batch's arrays)
                    self._apply_native_evidence(pending.pop(0))
            for s0, s1, seq in complex_ranges:
                self.cur_seq = seq
                lines = buf[off + s0:off + s1].decode("latin-1").splitlines(True)
                recs = [_parse_sam_line(ln, name2tid, 0) for ln in lines if ln.strip() and not ln.startswith("@")]
                nf0 = self.n_fragments
                for mate1, mate2 in iter_fragments(recs, N):
                    self.add_fragment(mate1, mate2)
                self.n_fragments = nf0 + 1

        def parse_piece(ing, buf, start, end, final, keep):
            """parse buf[start:end] with one handle; keep=False: hand every call's output straight to consume() (single
            thread), keep=True: return snapshots for the main thread.  Returns (results, offset reached)"""
            a, o = ing.a, ing.out
            off, out = start, []
            while off < end:
                used = ing.parse(buf, off, final, end)
                n, m = int(o.n_rows), int(o.n_frag_records)
                cx = [(int(a["cx_start"][k]), int(a["cx_end"][k]), int(a["cx_seq"][k])) for k in range(int(o.n_complex))]
                if keep:
                    out.append((off, n, m, int(o.max_l), int(o.n_fragments), [o.counters[k] for k in range(8)], cx, ing.snapshot(n, m)))
                else:
                    consume(buf, off, a, n, m, int(o.max_l), int(o.n_fragments), [o.counters[k] for k in range(8)], cx, ing.cap)
                off += used
                if used == 0:
                    break
                if n < ing.cap - 1 and len(cx) < int(o.cap_complex) and not final:
                    break  # (a fragment has up to two rows) the rest is an incomplete fragment: wait for the next chunk
            return out, off

        def consume_parsed(buf, jobs):
            for job in jobs:
                results, _ = job.result()
                for (o0, n, m, max_l, n_frag, counters, cx, arrays) in results:
                    consume(buf, o0, arrays, n, m, max_l, n_frag, counters, cx, n)

        try:
            if threads > 1:
                from concurrent.futures import ThreadPoolExecutor

                pool = ThreadPoolExecutor(threads + 1)
            # Three things overlap: the file read of chunk k+2 (one pool thread), the parse of chunk k+1 (parser threads, in
            # C++ without the GIL) and, on this thread, the GPU call + evidence rules + python-path fragments of chunk k.
            # The only dependency between chunks is the tail an earlier parse left over (an incomplete last fragment).
            ahead = pool.submit(read_chunk) if pool is not None else None
            waiting = None  # (buf, jobs) of the chunk whose results the main thread has not consumed yet
            T = self.t_ingest = {"read_wait": 0.0, "cut": 0.0, "consume": 0.0, "parse_wait": 0.0}  # where this thread's time goes
            clock = time.perf_counter
            carry = b""
            while not eof:
                t0 = clock()
                buf, got = ahead.result() if ahead is not None else read_chunk()
                T["read_wait"] += clock() - t0
                eof = got == 0  # (a stream may return short chunks before its end: BAM text comes in whole lines)
                if ahead is not None and not eof:
                    ahead = pool.submit(read_chunk)
                t0 = clock()
                if len(carry) <= HEAD:
                    lo, hi = HEAD - len(carry), HEAD + got
                    buf[lo:HEAD] = carry
                else:  # (a fragment of more than a megabyte of text)
                    buf = bytearray(carry) + buf[HEAD:HEAD + got]
                    lo, hi = 0, len(buf)
                if hi == lo:
                    break
                cuts = (_piece_cuts(buf, threads, lo, hi) if threads > 1 and hi - lo > int(getattr(opt, "ingest_piece_bytes", 4 << 20))
                        else [lo, hi])
                T["cut"] += clock() - t0
                if len(cuts) == 2:
                    t0 = clock()
                    if waiting is not None:
                        consume_parsed(*waiting)
                        waiting = None
                    _, off = parse_piece(ings[0], buf, lo, hi, eof, False)
                    T["consume"] += clock() - t0
                else:
                    t0 = clock()
                    while len(ings) < len(cuts) - 1:
                        ings.append(mk(0, False, piece_cap))
                    T["handles"] = T.get("handles", 0.0) + clock() - t0
                    # every piece numbers its fragments from its own base: more than it can hold apart (a SAM line is > 16 bytes)
                    stride = max(b - a0 for a0, b in zip(cuts, cuts[1:])) // 16 + 2
                    next_ord = ings[0].next_fragment()
                    jobs = []
                    for k in range(len(cuts) - 1):
                        if k:
                            ings[k].set_position(next_ord + k * stride, False)
                        last = k == len(cuts) - 2
                        jobs.append(pool.submit(parse_piece, ings[k], buf, cuts[k], cuts[k + 1], eof if last else True, True))
                    t0 = clock()
                    if waiting is not None:  # (while the pool parses this chunk)
                        consume_parsed(*waiting)
                    T["consume"] += clock() - t0
                    waiting = (buf, jobs)
                    t0 = clock()
                    off = jobs[-1].result()[1]
                    for job in jobs:
                        job.result()  # (every handle is free again; a parser error surfaces here)
                    T["parse_wait"] += clock() - t0
                    ings[0].set_position(next_ord + (len(cuts) - 1) * stride, False)
                carry = bytes(buf[off:hi])
                if eof and carry.strip():
                    raise ValueError("unparsable trailing SAM text")
            t0 = clock()
            if waiting is not None:
                consume_parsed(*waiting)
            for res in pending:
                self._apply_native_evidence(res)
            pending = []
            T["consume"] += clock() - t0
            self.flush()
        finally:
            t0 = time.perf_counter()
            if pool is not None:
                pool.shutdown()
            for ing in ings:
                ing.close()
            if getattr(self, "t_ingest", None) is not None:
                self.t_ingest["close"] = time.perf_counter() - t0

    def _native_batch(self, buf, off, a, n, m, max_l, n_words, plane_stride, lib, pool=None):
        """scan + record the n rows of m fragments the native ingest produced (on this thread: the context's stream takes one
        batch at a time), then their evidence (_native_evidence) -- on a pool thread when there is one and `a` is a snapshot:
        the rules run in C++ without the interpreter lock, beside the next batch's GPU call.  Returns what
        _apply_native_evidence takes (or a future of it); results are applied in stream order."""
        idx = a["frag_seq"][:n].astype(np.uint64) * np.uint64(64) + a["idx_k"][:n].astype(np.uint64)
        t0 = time.perf_counter()
        hits = self.eng.batch_host_planes(n, a["chrom"], a["a_start"], a["b_end"], a["l"], a["flags"], a["rlo"], a["rhi"], a["rn"],
                                          n_words, plane_stride, max(max_l, 0), a["wden"], a["q_a"], a["q_b"], a["read_hash"],
                                          a["qname_hash"], idx=idx, emit=True)
        self.t_scan += time.perf_counter() - t0
        self.n_pairs_scanned += n
        if pool is not None:
            return pool.submit(self._native_evidence, buf, off, a, n, m, hits, lib)
        return self._native_evidence(buf, off, a, n, m, hits, lib)

    def _native_evidence(self, buf, off, a, n, m, hits, lib):
        """the evidence rules of record_hits (find_circ.py:1276-1439) for all fragments of a batch at once in C++
        (fc_ingest_evidence: with at most two spans per fragment, back-splices first, every rule is a comparison between the
        fragment's columns) -- same outcome as _record_hits(), which stays the reading of the reference for everything else.
        Touches no state of the run: returns (counters, events | None, multi-event rows | None, reads | None)."""
        from .ingest import EV_LIN0, EV_LIN0_OUT, EV_LIN1, EV_LIN1_OUT, EV_UN, EV_UN_OUT, evidence, text_address

        opt = self.opt
        ev = evidence(lib, a, hits, n, m, off, opt.asize, FLAG_BIT)
        if not ev["any_hit"]:
            return ev["counters"], None, None, None
        # ---- per-junction flags (find_circ.py:1325-1327, 1433-1437)
        events = (ev["ev_key"], ev["ev_hash"], ev["ev_mask"]) if len(ev["ev_mask"]) else None
        base = text_address(buf)

        def gather(off3, len3):
            blob = np.empty(int(np.maximum(len3, 0).sum()), dtype=np.uint8)
            got = lib.fc_text_gather(base, len(len3), off3.ctypes.data, len3.ctypes.data, blob.ctypes.data)
            if got != len(blob):
                raise RuntimeError("fc_text_gather failed (%d)" % got)
            return blob

        # ---- multi-event rows (find_circ.py:1429-1431)
        multi = None
        if opt.multi_events:
            cls = ev["cls"]
            me = np.nonzero(cls & (EV_LIN0 | EV_LIN1 | EV_UN))[0]
            if len(me):
                txt_off = a["f_txt_off"][:6 * m].reshape(m, 2, 3)
                txt_len = a["f_txt_len"][:6 * m].reshape(m, 2, 3)
                off3 = np.ascontiguousarray(txt_off[me, 0] + off)
                len3 = np.ascontiguousarray(txt_len[me, 0])
                len3[:, 1:] = -1  # (the name only)
                c = cls[me]
                multi = dict(
                    seqs=a["f_seq"][:m][me].copy(), names=gather(off3, len3), name_len=len3[:, 0].copy(), ck=ev["ck"][me],
                    lin=[((c & EV_LIN0) > 0, (c & EV_LIN0_OUT) > 0, ev["key0"][me]), ((c & EV_LIN1) > 0, (c & EV_LIN1_OUT) > 0, ev["key1"][me])],
                    un=((c & EV_UN) > 0, (c & EV_UN_OUT) > 0, a["f_un_tid"][:m][me].copy(), a["f_un_pos"][:m][me].astype(np.int64),
                        a["f_un_aend"][:m][me].astype(np.int64)))
        # ---- the reads of every fragment with a junction (find_circ.py:1439, 1442-1447): name, sequence and qualities of
        # both mates go into one compact blob (C++); the FASTQ records are formatted from it when the junction names are
        # known (reads_text)
        reads = dict(seqs=ev["r_seq"], blob=gather(ev["r_off3"], ev["r_len3"]), len3=ev["r_len3"], k0=ev["r_k0"], k1=ev["r_k1"],
                     mask=ev["r_mask"])
        return ev["counters"], events, multi, reads

    def _apply_native_evidence(self, res):
        """what _native_evidence returned for one batch (or a future of it) into the run's state; call in stream order"""
        counters, events, multi, reads = res.result() if hasattr(res, "result") else res
        for key, cnt in zip(("circ_spliced", "circ_no_bp", "lin_spliced", "lin_no_bp"), counters):
            if cnt:  # the reference's counter dict only holds keys that were incremented (find_circ.py:1146)
                self.N[key] += float(cnt)
        if events is not None:
            self.ev_native.append(events)
        if multi is not None:
            self.native_multi.append(multi)
        if reads is not None:
            self.native_reads.append(reads)

    # ------------------------------------------------------------------ outputs
    def finalize(self, dist=None, torch_dev=None):
        """aggregate on the GPU and build the text outputs"""
        if dist is not None and dist.get_world_size() > 1:
            from . import parallel

            parallel.exchange_records(self.eng, dist, torch_dev, 0)
        nj = self.eng.agg_finalize(0)
        junc = self.eng.agg_fetch(nj)
        if dist is not None and dist.get_world_size() > 1:
            from . import parallel

            junc = parallel.gather_junctions(junc, dist, torch_dev)
        self.junctions = junc
        return junc

    def _names(self, junc) -> Dict[tuple, str]:
        """discovery-order names, counting junctions that are filtered from the output too (find_circ.py:683-686)"""
        cached = getattr(self, "_names_cache", None)
        if cached is not None and cached[0] is junc:
            return cached[1]
        names = {}
        counts = [0, 0]
        prefix = ("circ", "lin")
        cn = self.eng.chrom_names
        for r in junc:
            kind = (int(r["sk"]) >> 1) & 1
            key = (int(r["chrom"]), int(r["start"]), int(r["end"]), "-" if int(r["sk"]) & 1 else "+", kind)
            if key in self.known:
                names[key] = self.known[key]
                continue
            counts[kind] += 1
            names[key] = "%s_%s_%06d" % (self.opt.name, prefix[kind], counts[kind])
        self._names_cache = (junc, names)
        return names

    def categories(self, r, inf: Optional[JunctionInfo], min_dist, min_ov=None, min_nh=None, fields=None) -> List[str]:
        """Hit.categories (find_circ.py:601-654).  r: the junction's row, or None with fields = (signal text, best_q_left,
        best_q_right, n_uniq_bridges, end - start) and min_ov / min_nh given"""
        opt = self.opt
        cats = []
        if fields is None:
            fields = (decode_signal((int(r["sk"]) >> 16) & 0xFFF), int(r["best_q_left"]), int(r["best_q_right"]),
                      float(r["n_uniq_bridges"]), int(r["end"]) - int(r["start"]))
            min_ov = int(r["min_ov"]) if min_ov is None else min_ov
            min_nh = int(r["min_n_hits"]) if min_nh is None else min_nh
        signal, ql, qr, bridges, span = fields
        if signal != "GTAG":
            cats.append("NON_CANONICAL")
        if ql == 0 or qr == 0:
            cats.append("WARN_NON_UNIQUE_ANCHOR")
        if bridges == 0:
            cats.append("WARN_NO_UNIQ_BRIDGES")
        if min_nh > 1:
            cats.append("WARN_AMBIGUOUS_BP")
        ov, ed = min_ov, int(min_dist)
        if ov == 0 and ed == 0:
            pass
        elif ov < 2 and ed < 2:
            cats.append("WARN_EXT_1MM")
        elif ov >= 2 or ed >= 2:
            cats.append("WARN_EXT_2MM+")
        if span < opt.short_threshold:
            cats.append("SHORT")
        elif span > opt.huge_threshold:
            cats.append("HUGE")
        if inf is not None and inf.n_names:
            if not inf.unbroken:
                cats.append("WARN_ALWAYS_BROKEN")
            if not inf.unwarned:
                cats.append("WARN_ALWAYS_WARN")
        return cats

    def bed_text(self, kind: int) -> str:
        """store_list (find_circ.py:695-730)"""
        opt = self.opt
        J = self.junctions
        names = self._names(J)
        cn = self.eng.chrom_names
        info = self.junction_info()
        lines = ["#" + "\t".join(BED_HEADER) + "\n"]
        if len(J) == 0:
            return lines[0]
        # the filters of :706-717 on whole columns, the surviving rows as python tuples (a numpy record costs a microsecond per field)
        sk_all = J["sk"].astype(np.int64)
        keep = ((sk_all >> 1) & 1) == kind
        ql_all, qr_all = J["best_q_left"].astype(np.int64), J["best_q_right"].astype(np.int64)
        if opt.halfunique:
            keep &= ~((ql_all < opt.min_uniq_qual) & (qr_all < opt.min_uniq_qual))
        else:
            keep &= ~((ql_all < opt.min_uniq_qual) | (qr_all < opt.min_uniq_qual))
        if not opt.report_nobridges:
            keep &= J["n_uniq_bridges"] != 0
        rows = J[keep]
        cols = [rows[f].tolist() for f in ("chrom", "start", "end", "sk", "best_q_left", "best_q_right", "n_uniq_bridges", "n_weighted",
                                           "min_dist", "min_ov", "min_n_hits", "n_frags", "n_spanned", "n_uniq")]
        sig_text: Dict[int, str] = {}
        for chrom, start, end, sk, ql, qr, bridges, w, min_dist, min_ov, min_nh, n_frags, n_spanned, n_uniq in zip(*cols):
            strand = "-" if sk & 1 else "+"
            key = (chrom, start, end, strand, kind)
            inf = info.get(key)
            if inf is not None and inf.flags:
                flags = sorted(inf.flags)
                fcounts = [inf.flags[f] for f in flags]
            else:
                flags, fcounts = ["N/A"], [0]
            if key in self.known:  # the placeholder splice of a known site (find_circ.py:676)
                min_dist, min_ov, min_nh = min(min_dist, 10), min(min_ov, 10), 1
            sig = (sk >> 16) & 0xFFF
            if sig not in sig_text:
                sig_text[sig] = decode_signal(sig)
            cats = self.categories(None, inf, min_dist, min_ov, min_nh, (sig_text[sig], ql, qr, bridges, end - start))
            # with --max-mismatch 0 the reference's edit distance is a python bool (find_circ.py:868-870)
            edits = bool(min_dist) if opt.maxdist == 0 else min_dist
            ws = py2_str(w)
            lines.append("\t".join((
                cn[chrom], str(start), str(end), names[key], str(n_frags), strand, ws, str(n_spanned), str(n_uniq), py2_str(bridges),
                str(ql), str(qr), opt.name, ws, py2_str(edits), str(min_ov), str(min_nh), sig_text[sig], "N/A", ",".join(sorted(cats)),
                ",".join(flags), ",".join(str(c) for c in fcounts))) + "\n")
        return "".join(lines)

    def reads_text(self) -> str:
        """write_read (find_circ.py:1442-1447)"""
        names = self._names(self.junctions)
        out = []
        if self.explicit_idx:
            self.reads_out.sort(key=lambda e: e[0])  # native + python paths interleave: restore stream order (stable)
        for _, qname, seq, qual, keys, flags in self.reads_out:
            # (most reads support one junction and carry no flag: keep that case cheap)
            nm = names[keys[0]] if len(keys) == 1 else ",".join(sorted(names[k] for k in keys))
            head = qname + " " + nm + " " + (",".join(flags) if flags else "")
            out.append("@" + head + "\n" + seq + "\n+" + head + "\n" + str(qual) + "\n")
        if not self.native_reads:
            return "".join(out)
        # the reads of the native ingest: formatted per batch in C++, then merged with the others by stream position
        from . import _lib

        lib = _lib.load()
        py_seqs = np.array([e[0] for e in self.reads_out], dtype=np.int64)
        pieces, done = [], 0
        # the tail of a record's header ("<junction names> <flags>") is the same for every read of a (junctions, flags)
        # combination: one table of tails for the whole run, one index per read
        uniq, inverse = _unique_rows(np.concatenate([np.concatenate([b["k0"], b["k1"], b["mask"][:, None]], axis=1) for b in self.native_reads]))
        jn = []
        for u in uniq.tolist():
            nm = names[(u[0], u[1], u[2], "-" if u[3] else "+", u[4])]
            if u[5] >= 0:
                nm = ",".join(sorted((nm, names[(u[5], u[6], u[7], "-" if u[8] else "+", u[9])])))
            jn.append((nm + " " + flag_text(u[10])).encode("latin-1"))
        name_len = np.array([len(x) for x in jn], dtype=np.int32)
        name_off = np.zeros(len(jn), dtype=np.int64)
        name_off[1:] = np.cumsum(name_len[:-1])
        names_blob = b"".join(jn)
        inverse = np.ascontiguousarray(inverse.reshape(-1), dtype=np.int32)
        at = 0
        for b in self.native_reads:
            m = len(b["seqs"])
            name_idx = inverse[at:at + m]
            at += m
            rec_off = np.zeros(m + 1, dtype=np.int64)
            args = (b["blob"].ctypes.data, m, b["len3"].ctypes.data, name_idx.ctypes.data, names_blob, name_off.ctypes.data,
                    name_len.ctypes.data)
            need = lib.fc_fastq_format(*args, None, 0, rec_off.ctypes.data)
            text = np.empty(max(int(need), 1), dtype=np.uint8)
            got = lib.fc_fastq_format(*args, text.ctypes.data, int(need), rec_off.ctypes.data)
            if got != need:
                raise RuntimeError("fc_fastq_format failed (%d)" % got)
            raw = str(memoryview(text[:got]), "latin-1")  # (one copy: the bytes are ASCII)
            # python-path reads whose stream position falls inside this batch split it
            lo = 0
            hi_py = int(np.searchsorted(py_seqs, b["seqs"][-1], side="right")) if m else done
            for j in range(done, hi_py):
                cut = int(np.searchsorted(b["seqs"], py_seqs[j], side="left"))
                pieces.append(raw[int(rec_off[lo]):int(rec_off[cut])])
                pieces.append(out[j])
                lo = cut
            pieces.append(raw[int(rec_off[lo]):])
            done = max(done, hi_py)
        pieces.extend(out[done:])
        return pieces[0] if len(pieces) == 1 else "".join(pieces)

    def _multi_rows(self):
        """python-path and native rows together, in stream order"""
        rows = list(self.multi_out)
        cn = self.eng.chrom_names
        for b in self.native_multi:
            ends = np.cumsum(b["name_len"])
            text = b["names"].tobytes().decode("latin-1")
            ck = b["ck"].tolist()
            lin = [(ev.tolist(), out.tolist(), key.tolist()) for ev, out, key in b["lin"]]
            un_ev, un_out, un_tid, un_pos, un_aend = (x.tolist() for x in b["un"])
            for k, seq in enumerate(b["seqs"].tolist()):
                sets = (set(), set(), set(), set())  # lin cons, lin incons, mate cons, mate incons
                for ev, out, key in lin:
                    if ev[k]:
                        g, s0, e0, mi, _ = key[k]
                        sets[1 if out[k] else 0].add((cn[g], s0, e0, "-" if mi else "+"))
                if un_ev[k]:
                    sets[3 if un_out[k] else 2].add((self.sam_chroms[un_tid[k]], un_pos[k], un_aend[k], "*"))
                g, s0, e0, mi, kd = ck[k]
                rows.append((seq, text[int(ends[k]) - int(b["name_len"][k]):int(ends[k])], (g, s0, e0, "-" if mi else "+", kd)) + sets)
        if self.native_multi or self.explicit_idx:
            rows.sort(key=lambda e: e[0])
        return rows

    def multi_text(self) -> str:
        """MultiEventRecorder (find_circ.py:733-763)"""
        names = self._names(self.junctions)
        cn = self.eng.chrom_names
        lines = ["#" + "\t".join(MULTI_HEADER) + "\n"]
        for _, frag, ck, lin_cons, lin_incons, un_cons, un_incons in self._multi_rows():
            score = len(lin_cons) - 10 * len(lin_incons) + len(un_cons) - 10 * len(un_incons)
            gid, start, end, strand, _ = ck
            cols = [cn[gid], str(start), str(end), "ME:" + names[ck], str(score), strand, frag]
            cols.append(",".join("%d-%d" % (s, e) for c, s, e, _ in sorted(lin_cons)) if lin_cons else "NO_LIN_CONS")
            cols.append(",".join("[%s:%d-%d]" % (c, s, e) for c, s, e, _ in sorted(lin_incons)) if lin_incons else "NO_LIN_INCONS")
            cols.append(",".join("%d-%d" % (s, e) for c, s, e, _ in sorted(un_cons)) if un_cons else "NO_UNSPLICED_CONS")
            cols.append(",".join("[%s:%d-%d]" % (c, s, e) for c, s, e, _ in sorted(un_incons)) if un_incons else "NO_UNSPLICED_INCONS")
            lines.append("\t".join(cols) + "\n")
        return "".join(lines)

    def test_text(self) -> str:
        """test_results.tsv (--test), rows in stream order"""
        return "".join(row + "\n" for _, row in sorted(self.test_out, key=lambda e: e[0]))

    def counters_text(self) -> str:
        """the N dump of find_circ.py:1605-1607"""
        return "".join("%s=%s\n" % (k, py2_str(float(self.N[k]))) for k in sorted(self.N))

    def close(self):
        if self.own_engine:
            self.eng.close()
