"""
bench_workload.py -- the synthetic inputs of bench.py at the sizes BASELINE.json names, generated
with torch on whatever device is at hand.

Everything is a pure function of (seed, position) or (seed, pair index) through a 64-bit mixing function evaluated with
wrap-around int64 arithmetic, so a CUDA device and the CPU produce the same bytes: the GPU arm materialises the whole
genome and its shard of the pairs on the device, the CPU arm (oracle) evaluates only the stretches of the genome it
touches (LazyGenome) and a prefix of the same pairs -- the CPU sample IS a prefix of the GPU workload.

  genome    chromosomes of i.i.d. uniform ACGT, `n_frac` of the bases in N runs, no soft-masking at bench sizes
  junctions `n_circ` back-splices + `n_lin` linear junctions flanked by GT/AG (CT/AC on '-'), planted like
            find_circ2_b200.synth.plant_junctions (find_circ.py:924-954 is what recognises them)
  pairs     two-segment reads across the junctions, the mix of find_circ2_b200.synth.make_pairs: Zipf popularity, decoys,
            chromosome-edge pairs, non-unique anchors, substitutions, a few N; optionally as mate pairs that share a name

Bench tooling, not product code: nothing in find_circ2_b200/ imports it.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np
import torch

# hg19 chr1..22, X, Y (test_data/test_norm.sam:1-93 of the reference lists the same @SQ lengths)
HG19_SIZES = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022, 141213431, 135534747,
              135006516, 133851895, 115169878, 107349540, 102531392, 90354753, 81195210, 78077248, 59128983, 63025520,
              48129895, 51304566, 155270560, 59373566]

_M1 = 0xBF58476D1CE4E5B9 - (1 << 64)
_M2 = 0x94D049BB133111EB - (1 << 64)
_G1 = 0x9E3779B97F4A7C15 - (1 << 64)
_G2 = 0xD1B54A32D192ED03 - (1 << 64)


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (two's complement wrap-around = arithmetic mod 2^64)"""
    x = x ^ _lsr(x, 30)
    x = x * _M1
    x = x ^ _lsr(x, 27)
    x = x * _M2
    return x ^ _lsr(x, 31)


def rnd(idx: torch.Tensor, stream: int, seed: int) -> torch.Tensor:
    """64 random bits per element of `idx` (int64), independent per (stream, seed)"""
    return mix64(idx * _G1 + ((stream * _G2 + seed * 0x632BE59BD9B4E019) % (1 << 63)))


def uniform(idx, stream, seed) -> torch.Tensor:
    """float64 in [0, 1)"""
    return _lsr(rnd(idx, stream, seed), 11).to(torch.float64) * (1.0 / 9007199254740992.0)


def randint(idx, stream, seed, lo: int, hi: int) -> torch.Tensor:
    """int64 in [lo, hi)"""
    return lo + (_lsr(rnd(idx, stream, seed), 1) % (hi - lo))


_ACGT = torch.tensor(list(b"ACGT"), dtype=torch.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


@dataclasses.dataclass
class Config:
    name: str
    sizes: List[int]
    n_circ: int
    n_lin: int
    n_pairs: int              # whole job
    read_len: int = 100
    asize: int = 20
    margin: int = 2
    maxdist: int = 2
    error_rate: float = 0.005
    zipf: float = 1.0
    paired: bool = False      # pairs 2f and 2f+1 are the mates of fragment f (one read name)
    p_same: float = 0.5       # paired: the second mate crosses the junction of the first
    frac_decoy: float = 0.10
    frac_nonuniq: float = 0.02
    frac_edge: float = 0.01
    frac_inner_shift: float = 0.05
    frac_read_n: float = 0.005
    frac_no_xs: float = 0.05
    frac_ambiguous: float = 0.0  # reserved
    n_frac: float = 0.005
    seed: int = 1
    scaling: str = "strong"   # how n_pairs relates to the GPU count: "strong" = whole job fixed, "weak" = per GPU
    min_uniq: int = 2         # the is_uniq pre-filter of find_circ.py:1299-1301 (0: keep everything, the fmt-1.2 behaviour)
    halfunique: bool = False        # --halfuniq: keep junctions with one uniquely placed side (find_circ.py:706-713)
    report_nobridges: bool = False  # --report_nobridge: keep junctions without a unique bridge (find_circ.py:715-717)
    workload: str = ""


def configs() -> Dict[str, Config]:
    """the configurations of BASELINE.json (configs[1..4]; configs[0] is the reference's CDR1as test, a parity case)"""
    return {
        "2": Config("2", [5000000] * 20, 10000, 500, 1000000, scaling="weak",
                    workload="configs[1]: synthetic 100 Mb genome (20 chrom, 0.5% N), 10k planted circRNAs, 1M anchor pairs per GPU from 100-nt reads, a=20 m=2 d=2, 0.5% substitutions, 10% decoys"),
        "3": Config("3", HG19_SIZES, 100000, 5000, 50000000, paired=True,
                    workload="configs[2]: hg19-sized synthetic genome (3.1 Gb, 24 chrom, 0.5% N, replicated per GPU), 100k planted circRNAs, 50M anchor pairs in the whole job from 2x100-nt mate pairs (mates share a name), a=20 m=2 d=2, 0.5% substitutions, 10% decoys"),
        "4": Config("4", [5000000] * 20, 10000, 500, 4000000, read_len=150, error_rate=0.02, scaling="weak", min_uniq=0, frac_nonuniq=0.10,
                    halfunique=True, report_nobridges=True,
                    workload="configs[3]: 100 Mb genome, 150-nt reads (l=114), 1/2/3% substitutions, no per-span uniqueness filter (--halfuniq --report_nobridge semantics: one-sided anchors and zero-bridge junctions stay in), 4M anchor pairs per GPU"),
        "5": Config("5", HG19_SIZES, 10000000, 100000, 500000000 // 8, zipf=0.6, scaling="weak",
                    workload="configs[4]: 500M anchor pairs over 8 GPUs = 62.5M per GPU, 10M distinct planted junctions (Zipf 0.6 tail: most junctions carry 1-5 reads), hg19-sized genome"),
    }


# ------------------------------------------------------------------------------------------------ genome
class Spec(object):
    """genome + junctions of a Config as pure functions of position (no base array held)"""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.sizes = np.asarray(cfg.sizes, dtype=np.int64)
        self.names = ["chr%d" % (k + 1) for k in range(len(self.sizes))]
        self.off = np.zeros(len(self.sizes) + 1, dtype=np.int64)  # flat (unpadded) coordinate of base 0 of each chromosome
        self.off[1:] = np.cumsum(self.sizes)
        self.total = int(self.off[-1])
        seed = cfg.seed
        # N runs: k per chromosome, uniform start, uniform length in [50, 5000]
        runs_s, runs_e = [], []
        for c, size in enumerate(self.sizes.tolist()):
            k = int(cfg.n_frac * size / 2525.0)
            if k <= 0 or size < 20000:
                continue
            idx = torch.arange(k, dtype=torch.int64) + (c << 32)
            st = randint(idx, 101, seed, 0, size - 5000).numpy()
            ln = randint(idx, 102, seed, 50, 5001).numpy()
            runs_s.append(st + self.off[c])
            runs_e.append(st + ln + self.off[c])
        if runs_s:
            s, e = np.concatenate(runs_s), np.concatenate(runs_e)
            o = np.argsort(s, kind="stable")
            s, e = s[o], e[o]
            # merge overlapping runs so that membership is one searchsorted
            keep_s, keep_e = [], []
            cur_s, cur_e = int(s[0]), int(e[0])
            for a, b in zip(s[1:].tolist(), e[1:].tolist()):
                if a <= cur_e:
                    cur_e = max(cur_e, b)
                else:
                    keep_s.append(cur_s)
                    keep_e.append(cur_e)
                    cur_s, cur_e = a, b
            keep_s.append(cur_s)
            keep_e.append(cur_e)
            self.n_start, self.n_end = np.asarray(keep_s, np.int64), np.asarray(keep_e, np.int64)
        else:
            self.n_start = self.n_end = np.zeros(0, np.int64)
        self._plant()

    def _plant(self):
        """junction table + the bases the splice signals overwrite (find_circ2_b200.synth.plant_junctions semantics)"""
        cfg, seed = self.cfg, self.cfg.seed
        n = cfg.n_circ + cfg.n_lin
        idx = torch.arange(n, dtype=torch.int64)
        cum = np.cumsum(self.sizes / self.sizes.sum())
        chrom = np.minimum(np.searchsorted(cum, uniform(idx, 201, seed).numpy(), side="right"), len(self.sizes) - 1).astype(np.int64)
        lo, hi, margin = 200, 50000, 400
        sp = np.exp(np.log(lo) + uniform(idx, 202, seed).numpy() * (np.log(hi) - np.log(lo))).astype(np.int64)
        csz = self.sizes[chrom]
        sp = np.minimum(sp, np.maximum(csz - 2 * margin - 4, 8))
        start = np.maximum((margin + uniform(idx, 203, seed).numpy() * (csz - sp - 2 * margin)).astype(np.int64), 2)
        end = start + sp
        minus = uniform(idx, 204, seed).numpy() < 0.5
        circ = np.zeros(n, dtype=bool)
        circ[:cfg.n_circ] = True
        self.j_chrom, self.j_start, self.j_end, self.j_minus, self.j_circ = chrom, start, end, minus, circ
        # overwritten bases: positions (flat) and letters, junction order = write order (a later junction wins)
        base = self.off[chrom]
        p0 = np.where(circ, start - 2, start) + base  # left site, 2 bases
        p1 = np.where(circ, end, end - 2) + base      # right site, 2 bases
        left = np.where(circ[:, None], np.where(minus[:, None], np.frombuffer(b"AC", np.uint8), np.frombuffer(b"AG", np.uint8)),
                        np.where(minus[:, None], np.frombuffer(b"CT", np.uint8), np.frombuffer(b"GT", np.uint8)))
        right = np.where(circ[:, None], np.where(minus[:, None], np.frombuffer(b"CT", np.uint8), np.frombuffer(b"GT", np.uint8)),
                         np.where(minus[:, None], np.frombuffer(b"AC", np.uint8), np.frombuffer(b"AG", np.uint8)))
        pos = np.stack([p0, p0 + 1, p1, p1 + 1], axis=1).reshape(-1)
        val = np.concatenate([left, right], axis=1).reshape(-1).astype(np.uint8)
        o = np.argsort(pos, kind="stable")
        pos, val = pos[o], val[o]
        last = np.ones(len(pos), dtype=bool)
        last[:-1] = pos[1:] != pos[:-1]  # the last write to a position stays
        self.ov_pos, self.ov_val = pos[last], val[last]
        # Zipf popularity over a hashed permutation of the junctions; circ and linear drawn from one law
        w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), cfg.zipf)
        perm = np.argsort(rnd(idx, 205, seed).numpy(), kind="stable")
        ww = np.empty(n, dtype=np.float64)
        ww[perm] = w
        self.j_cum = np.cumsum(ww / ww.sum())

    # ---- bases
    def bases(self, flat_start: int, flat_end: int, device="cpu") -> torch.Tensor:
        """uint8 ASCII of flat positions [flat_start, flat_end)"""
        p = torch.arange(flat_start, flat_end, dtype=torch.int64, device=device)
        b = _ACGT.to(device)[(rnd(p, 1, self.cfg.seed) & 3)]
        del p
        lo = int(np.searchsorted(self.n_end, flat_start, side="right"))
        hi = int(np.searchsorted(self.n_start, flat_end, side="left"))
        for s, e in zip(self.n_start[lo:hi].tolist(), self.n_end[lo:hi].tolist()):
            b[max(s, flat_start) - flat_start:min(e, flat_end) - flat_start] = ord("N")
        lo = int(np.searchsorted(self.ov_pos, flat_start, side="left"))
        hi = int(np.searchsorted(self.ov_pos, flat_end, side="left"))
        if hi > lo:
            b[torch.from_numpy(self.ov_pos[lo:hi] - flat_start).to(device)] = torch.from_numpy(self.ov_val[lo:hi]).to(device)
        return b

    def materialize(self, device, chunk: int = 1 << 28) -> torch.Tensor:
        """the whole genome as one flat uint8 tensor on `device` (chromosomes back to back, no padding)"""
        out = torch.empty(self.total, dtype=torch.uint8, device=device)
        for s in range(0, self.total, chunk):
            e = min(self.total, s + chunk)
            out[s:e] = self.bases(s, e, device)
        return out

    def chrom_arrays(self, flat: torch.Tensor) -> List[np.ndarray]:
        """per-chromosome host views of a flat genome tensor (for fc_genome_load_ascii)"""
        host = flat.cpu().numpy() if flat.is_cuda else flat.numpy()
        return [host[self.off[c]:self.off[c + 1]] for c in range(len(self.sizes))]


class LazySeq(object):
    """str-like chromosome for the oracle (len + slicing), bases evaluated on demand"""

    def __init__(self, spec: Spec, c: int):
        self.spec, self.c = spec, c
        self.n = int(spec.sizes[c])

    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        start, stop, step = sl.indices(self.n)
        assert step == 1
        if stop <= start:
            return ""
        o = int(self.spec.off[self.c])
        return self.spec.bases(o + start, o + stop).numpy().tobytes().decode()


class SparseSeq(object):
    """str-like chromosome holding only the stretches the sample's reads and windows touch (what an mmap slice costs the
    reference is what a slice costs here); anything else falls back to LazySeq"""

    def __init__(self, spec: Spec, c: int, starts: np.ndarray, ends: np.ndarray):
        import bisect

        self._bisect = bisect.bisect_right
        self.lazy = LazySeq(spec, c)
        self.n = self.lazy.n
        o = int(spec.off[c])
        self.starts = starts.tolist()
        self.ends = ends.tolist()
        lens = ends - starts
        self.at = np.concatenate([[0], np.cumsum(lens)]).tolist()
        if len(starts):
            gpos = torch.from_numpy(np.concatenate([np.arange(a, b, dtype=np.int64) for a, b in zip(starts.tolist(), ends.tolist())]) + o)
            self.blob = _bases_at(spec, gpos).numpy().tobytes().decode()
        else:
            self.blob = ""

    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        start, stop, step = sl.indices(self.n)
        if stop <= start:
            return ""
        k = self._bisect(self.starts, start) - 1
        if k >= 0 and stop <= self.ends[k]:
            a = self.at[k] + (start - self.starts[k])
            return self.blob[a:a + (stop - start)]
        return self.lazy[sl]


def sparse_genome(spec: Spec, cols: Dict[str, torch.Tensor], pad: int = 300) -> Dict[str, SparseSeq]:
    """chromosome name -> SparseSeq covering [pos - pad, pos + read_len + pad) around both segments of every pair in cols"""
    chrom = cols["chrom"].cpu().numpy().astype(np.int64)
    R = spec.cfg.read_len
    out = {}
    for c, name in enumerate(spec.names):
        m = chrom == c
        p = np.concatenate([cols["a_pos"].cpu().numpy()[m], cols["b_pos"].cpu().numpy()[m]])
        if len(p) == 0:
            out[name] = SparseSeq(spec, c, np.zeros(0, np.int64), np.zeros(0, np.int64))
            continue
        size = int(spec.sizes[c])
        s = np.clip(p - pad, 0, size)
        e = np.clip(p + R + pad, 0, size)
        o = np.argsort(s, kind="stable")
        s, e = s[o], e[o]
        # merge overlapping stretches
        run_end = np.maximum.accumulate(e)
        new = np.ones(len(s), dtype=bool)
        new[1:] = s[1:] > run_end[:-1]
        first = np.nonzero(new)[0]
        last = np.concatenate([first[1:], [len(s)]]) - 1
        out[name] = SparseSeq(spec, c, s[first], run_end[last])
    return out


# ------------------------------------------------------------------------------------------------ pairs
def make_pairs(spec: Spec, i0: int, i1: int, device="cpu", flat: Optional[torch.Tensor] = None, chunk: int = 1 << 21,
               error_rate: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """pairs [i0, i1) of the job as the columns of find_circ2_b200.synth.PairTable (tensors on `device`).
    `flat`: materialised genome on the same device (else bases are evaluated per read)."""
    cfg, seed = spec.cfg, spec.cfg.seed
    R, asize = cfg.read_len, cfg.asize
    err = cfg.error_rate if error_rate is None else error_rate
    dev = torch.device(device)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    j_cum, j_chrom, j_start, j_end, j_circ = t(spec.j_cum), t(spec.j_chrom), t(spec.j_start), t(spec.j_end), t(spec.j_circ)
    sizes, off = t(spec.sizes), t(spec.off)
    nj = len(spec.j_cum)
    cols: Dict[str, List[torch.Tensor]] = {}

    def put(name, v):
        cols.setdefault(name, []).append(v)

    min_anchor = max(asize - 2, 1)
    for c0 in range(i0, i1, chunk):
        c1 = min(i1, c0 + chunk)
        i = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        n = c1 - c0
        # which junction: own draw, or (second mate) the junction of the first mate
        draw = i
        if cfg.paired:
            second = (i & 1) == 1
            same = second & (uniform(i >> 1, 11, seed) < cfg.p_same)
            draw = torch.where(same, i - 1, i)
        jx = torch.clamp(torch.searchsorted(j_cum, uniform(draw, 12, seed), right=True), max=nj - 1)
        chrom = j_chrom[jx]
        circ = j_circ[jx]
        j = randint(i, 13, seed, min_anchor, R - min_anchor + 1)
        a_end = torch.where(circ, j_end[jx], j_start[jx])
        b_pos = torch.where(circ, j_start[jx], j_end[jx])
        a_pos = a_end - j
        a_len = j
        b_len = R - j
        junc = jx.clone()
        # decoys: same orientation, second locus without a splice signal
        decoy = uniform(i, 14, seed) < cfg.frac_decoy
        span = randint(i, 15, seed, 300, 20000)
        back = uniform(i, 16, seed) < 0.7
        b_pos = torch.where(decoy, torch.where(back, a_pos - span, a_pos + span + R), b_pos)
        junc = torch.where(decoy, torch.full_like(junc, -1), junc)
        # pairs hugging a chromosome end: the windows run into the N padding (find_circ.py:194-211)
        edge = uniform(i, 17, seed) < cfg.frac_edge
        csz = sizes[chrom]
        at_end = uniform(i, 18, seed) < 0.5
        offs = randint(i, 19, seed, 0, 40)
        new_a_end = torch.where(at_end, csz - offs, 600 + offs)
        a_pos = torch.where(edge, new_a_end - j, a_pos)
        b_pos = torch.where(edge, torch.where(at_end, new_a_end - 500, offs), b_pos)
        junc = torch.where(edge, torch.full_like(junc, -1), junc)
        a_pos = torch.minimum(torch.clamp(a_pos, min=0), csz - a_len)
        b_pos = torch.minimum(torch.clamp(b_pos, min=0), csz - b_len - 8)
        # read = genome[a_pos : a_pos + a_len] + genome[b_pos : b_pos + b_len]
        col = torch.arange(R, dtype=torch.int64, device=dev)[None, :]
        gpos = torch.where(col < a_len[:, None], a_pos[:, None] + col, b_pos[:, None] + (col - a_len[:, None])) + off[chrom][:, None]
        if flat is not None:
            reads = flat[gpos]
        else:
            reads = _bases_at(spec, gpos)
        del gpos
        if err > 0:
            e_idx = i[:, None] * R + col
            hit = uniform(e_idx, 20, seed) < err
            shift = (1 + (_lsr(rnd(e_idx, 21, seed), 3) % 3)).to(torch.uint8)
            del e_idx
            code = torch.full_like(reads, 4)
            for k, ch in enumerate(b"ACGT"):
                code[reads == ch] = k
            new = _ACGT.to(dev)[((code + shift) & 3).to(torch.int64)]
            reads = torch.where(hit & (code < 4), new, reads)
            del code, new, hit, shift
        if cfg.frac_read_n > 0:
            rn = uniform(i, 22, seed) < cfg.frac_read_n
            at = randint(i, 23, seed, 0, R)
            rows = torch.nonzero(rn)[:, 0]
            reads[rows, at[rows]] = ord("N")
        # aligner over-extension: the inner boundary of the segments is not the true breakpoint
        shift = torch.where(uniform(i, 24, seed) < cfg.frac_inner_shift, randint(i, 25, seed, -4, 5), torch.zeros_like(i))
        shift = torch.where(a_len < asize, asize - a_len, shift)
        shift = torch.where(b_len < asize, -(asize - b_len), shift)
        shift = torch.maximum(shift, -b_pos)
        a_len2 = a_len + shift
        b_len2 = R - a_len2
        b_pos2 = b_pos + shift
        reverse = uniform(i, 26, seed) < 0.5
        as_a, as_b = a_len2.clone(), b_len2.clone()
        xs_a = torch.clamp(as_a - randint(i, 27, seed, 2, 30), min=0)
        xs_b = torch.clamp(as_b - randint(i, 28, seed, 2, 30), min=0)
        nonu = uniform(i, 29, seed) < cfg.frac_nonuniq
        side = uniform(i, 30, seed) < 0.5
        xs_a = torch.where(nonu & side, as_a - randint(i, 31, seed, 0, 2), xs_a)
        xs_b = torch.where(nonu & ~side, as_b - randint(i, 32, seed, 0, 2), xs_b)
        noxs = uniform(i, 33, seed) < cfg.frac_no_xs
        xs_a = torch.where(noxs, torch.full_like(xs_a, -1), xs_a)
        xs_b = torch.where(noxs & (uniform(i, 34, seed) < 0.5), torch.full_like(xs_b, -1), xs_b)
        for name, v in (("junc", junc), ("chrom", chrom.to(torch.int32)), ("a_pos", a_pos), ("a_len", a_len2.to(torch.int32)),
                        ("b_pos", b_pos2), ("b_len", b_len2.to(torch.int32)), ("reverse", reverse), ("reads", reads),
                        ("as_a", as_a.to(torch.int32)), ("xs_a", xs_a.to(torch.int32)), ("as_b", as_b.to(torch.int32)),
                        ("xs_b", xs_b.to(torch.int32)), ("primary_is_b", b_len2 > a_len2),
                        ("name_id", (i >> 1) if cfg.paired else i)):
            put(name, v)
    return {k: (torch.cat(v) if len(v) > 1 else v[0]) for k, v in cols.items()}


def _bases_at(spec: Spec, gpos: torch.Tensor) -> torch.Tensor:
    """bases at arbitrary flat positions without a materialised genome (CPU arm): hash + N runs + planted signals"""
    dev = gpos.device
    b = _ACGT.to(dev)[(rnd(gpos, 1, spec.cfg.seed) & 3)]
    if len(spec.n_start):
        ns, ne = torch.from_numpy(spec.n_start).to(dev), torch.from_numpy(spec.n_end).to(dev)
        k = torch.searchsorted(ns, gpos, right=True) - 1
        inside = (k >= 0) & (gpos < ne[torch.clamp(k, min=0)])
        b = torch.where(inside, torch.full_like(b, ord("N")), b)
    if len(spec.ov_pos):
        op, ovv = torch.from_numpy(spec.ov_pos).to(dev), torch.from_numpy(spec.ov_val).to(dev)
        k = torch.clamp(torch.searchsorted(op, gpos), max=len(spec.ov_pos) - 1)
        b = torch.where(op[k] == gpos, ovv[k], b)
    return b


def pair_table(cols: Dict[str, torch.Tensor]):
    """host PairTable (find_circ2_b200.synth) of a (small) column dict: what the SAM writers and the oracle take"""
    from find_circ2_b200 import synth

    h = {k: v.cpu().numpy() for k, v in cols.items()}
    return synth.PairTable(junc=h["junc"], chrom=h["chrom"], a_pos=h["a_pos"], a_len=h["a_len"], b_pos=h["b_pos"], b_len=h["b_len"],
                           reverse=h["reverse"], reads=h["reads"], as_a=h["as_a"], xs_a=h["xs_a"], as_b=h["as_b"], xs_b=h["xs_b"],
                           primary_is_b=h["primary_is_b"], name_id=h["name_id"])


def soa(cols: Dict[str, torch.Tensor], cfg: Config) -> Dict[str, torch.Tensor]:
    """what ingest produces for 2-segment reads (find_circ.py:1058-1140, 821-848): scan inputs + aggregation payload, minus
    the read / name hashes (the caller computes them with the library).  Rows failing the is_uniq filter are dropped
    before the scan, as find_circ.py:1299-1301 does; `row` keeps their position in the input."""
    eff = cfg.asize - cfg.margin
    R = cfg.read_len
    as_a, xs_a, as_b, xs_b = (cols[k].to(torch.int64) for k in ("as_a", "xs_a", "as_b", "xs_b"))
    uniq_a = torch.where(xs_a >= 0, as_a - xs_a, as_a)
    uniq_b = torch.where(xs_b >= 0, as_b - xs_b, as_b)
    keep = torch.minimum(uniq_a, uniq_b) >= cfg.min_uniq
    row = torch.nonzero(keep)[:, 0]
    a_pos, b_pos = cols["a_pos"][row], cols["b_pos"][row]
    a_len, b_len = cols["a_len"][row].to(torch.int64), cols["b_len"][row].to(torch.int64)
    n = len(row)
    dev = row.device
    return dict(
        row=row,
        chrom=cols["chrom"][row].to(torch.int32),
        a_start=(a_pos + eff).to(torch.int32),
        b_end=(b_pos + b_len - eff).to(torch.int32),
        l=torch.full((n,), R - 2 * eff, dtype=torch.int32, device=dev),
        flags=(((b_pos - (a_pos + a_len)) < 0).to(torch.uint8) | (cols["reverse"][row].to(torch.uint8) << 1)),
        internal=cols["reads"][row][:, eff:R - eff].contiguous(),
        reads=cols["reads"][row],
        wden=torch.ones(n, dtype=torch.uint8, device=dev),
        q_a=(as_a[row] - torch.clamp(xs_a[row], min=0)).to(torch.int16),
        q_b=(as_b[row] - torch.clamp(xs_b[row], min=0)).to(torch.int16),
        name_id=cols["name_id"][row],
    )
