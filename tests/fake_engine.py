"""
TEST DOUBLE for find_circ2_b200.engine.Engine -- lets the CPU-only test-suite exercise the HOST logic of the product
(fragment grouping, span batching, record_hits control flow, writers) against the reference goldens.
The scan is answered by the device code compiled for the host (tests/tools/scan_host_harness.cpp) and by the oracle
for tie lists; the aggregation by a python dict.  Never imported by the product.
"""
import numpy as np

import helpers as H
from find_circ2_b200 import _lib
from find_circ2_b200._lib import HIT_DTYPE, JREC_DTYPE, JUNCTION_DTYPE
from oracle import find_circ_oracle as O


class _G:
    def __init__(self, names, seqs):
        self.names, self.seqs = names, seqs


class FakeEngine:
    def __init__(self, device=0, asize=15, margin=2, maxdist=2, noncanonical=False, strandpref=False):
        self.lib = _lib.load()
        self.p = dict(asize=asize, margin=margin, maxdist=maxdist, noncanonical=int(noncanonical), strandpref=int(strandpref))
        self.chrom_names, self.chrom_sizes = [], []
        self.recs = []
        self.last = None

    def load_genome_fasta(self, path):
        og = O.Genome(path)
        self.og = og
        self.chrom_names = list(og.names)
        self.chrom_sizes = [og.size(n) for n in og.names]
        self._chrom_ids = {n: i for i, n in enumerate(self.chrom_names)}
        self.g = _G(self.chrom_names, [np.frombuffer(og.seqs[n].encode(), dtype=np.uint8).copy() for n in og.names])

    def chrom_id(self, name):
        return self.chrom_names.index(name) if name in self.chrom_names else (_ for _ in ()).throw(KeyError(name))

    def hash_read(self, seq):
        return int(self.lib.fc_hash_read(seq, len(seq), None))

    def hash_bytes(self, b):
        return int(self.lib.fc_hash_bytes(b, len(b)))

    def agg_reset(self):
        self.recs = []

    def batch_host_planes(self, n, chrom, a_start, b_end, l, flags, rlo, rhi, rn, n_words, plane_stride, max_l, wden, q_a, q_b,
                          read_hash, qname_hash, idx=None, idx_base=0, emit=True, out=None):
        """planes back to ASCII (test double only), then the ASCII path"""
        lmax = max(int(max_l), 1)
        internal = np.zeros((n, lmax), dtype=np.uint8)
        for i in range(n):
            for j in range(max(int(l[i]), 0)):
                w, b = j >> 5, j & 31
                if (int(rn[w * plane_stride + i]) >> b) & 1:
                    internal[i, j] = ord("N")
                else:
                    code = ((int(rlo[w * plane_stride + i]) >> b) & 1) | (((int(rhi[w * plane_stride + i]) >> b) & 1) << 1)
                    internal[i, j] = b"ACGT"[code]
        fl = flags[:n].copy() & 0xFB
        return self.batch_host(chrom[:n].copy(), a_start[:n].copy(), b_end[:n].copy(), l[:n].copy(), fl, internal, wden[:n].copy(),
                               q_a[:n].copy(), q_b[:n].copy(), read_hash[:n].copy(), qname_hash[:n].copy(), idx_base, emit=emit,
                               out=out, idx=idx)

    def batch_host(self, chrom, a_start, b_end, l, flags, internal, wden, q_a, q_b, read_hash, qname_hash, idx_base, emit=True,
                   out=None, want_hits=True, idx=None):
        p = self.p
        hits = H.harness_scan(self.g, chrom, a_start, b_end, l, flags, internal, p["margin"], p["maxdist"], p["noncanonical"],
                              p["strandpref"])
        self.last = dict(chrom=chrom, a_start=a_start, b_end=b_end, l=l, flags=flags, internal=internal, wden=wden, q_a=q_a,
                         q_b=q_b, rh=read_hash, qh=qname_hash, hits=hits, idx=idx)
        if emit:
            self.batch_emit(None, idx_base)
        res = hits.view(HIT_DTYPE).reshape(-1)
        if out is not None:
            out[:] = res
            return out
        return res

    def batch_emit(self, mask, idx_base):
        L = self.last
        for i, h in enumerate(L["hits"]):
            nh = int(h[2]) & 0xFFFF
            if nh == 0 or (mask is not None and not mask[i]):
                continue
            back = int(L["flags"][i]) & 1
            rh = int(L["rh"][i])
            r = np.zeros((), dtype=JREC_DTYPE)
            r["chrom"], r["start"], r["end"] = L["chrom"][i], np.int32(np.uint32(h[0]).astype(np.int32)), np.uint32(h[1]).astype(np.int32)
            r["sk"] = (int(h[3]) & 1) | (0 if back else 2) | ((rh & 1) << 2) | (int(L["wden"][i]) << 8) | (((int(h[3]) >> 1) & 0xFFF) << 16)
            r["idx"] = idx_base + i if L.get("idx") is None else int(L["idx"][i])
            r["read_hash"], r["qname_hash"] = rh, int(L["qh"][i])
            r["q_left"], r["q_right"] = (L["q_b"][i], L["q_a"][i]) if back else (L["q_a"][i], L["q_b"][i])
            r["n_hits"], r["dist"], r["ov"] = nh, (int(h[2]) >> 16) & 0xFF, int(h[2]) >> 24
            self.recs.append(r)

    def batch_ties(self, n_hits):
        L = self.last
        p = self.p
        opt = O.Options(asize=p["asize"], margin=p["margin"], maxdist=p["maxdist"], noncanonical=bool(p["noncanonical"]),
                        strandpref=bool(p["strandpref"]), allhits=True)
        off = np.zeros(len(n_hits) + 1, dtype=np.int64)
        off[1:] = np.cumsum(n_hits)
        ties = np.zeros(int(off[-1]), dtype=HIT_DTYPE)
        for i in range(len(n_hits)):
            if n_hits[i] == 0:
                continue
            li = int(L["l"][i])
            c = self.chrom_names[int(L["chrom"][i])]
            a0, b1 = int(L["a_start"][i]), int(L["b_end"][i])
            fl = int(L["flags"][i])
            hs = O.scan_windows(self.og.get(c, a0, a0 + li + 2).upper(), self.og.get(c, b1 - li - 2, b1).upper(),
                                L["internal"][i, :li].tobytes().decode().upper(), c, a0, b1, bool(fl & 1), "-" if fl & 2 else "+", opt)
            assert len(hs) == n_hits[i]
            for k, h in enumerate(hs):
                sig = sum("ACGTN".index(ch) << (3 * q) for q, ch in enumerate(h.gtag))
                ties[off[i] + k] = (h.start, h.end, h.n_hits | (int(h.dist) << 16) | (h.ov << 24), (1 if h.strand == "-" else 0) | (sig << 1))
        return off, ties

    def agg_append_host(self, recs):
        for r in recs:
            self.recs.append(r.copy())

    # ---- multi-rank exchange (find_circ2_b200.parallel.exchange_records): partition by a key hash, replace by what arrives
    def agg_n_records(self):
        return len(self.recs)

    def agg_partition(self, world, send, stream=0):
        import torch

        recs = np.array(self.recs, dtype=JREC_DTYPE) if self.recs else np.zeros(0, dtype=JREC_DTYPE)
        k = (recs["chrom"].astype(np.uint64) * np.uint64(1000003) + recs["start"].astype(np.int64).astype(np.uint64) * np.uint64(10007)
             + recs["end"].astype(np.int64).astype(np.uint64) * np.uint64(101) + (recs["sk"] & 3).astype(np.uint64))
        d = (k % np.uint64(world)).astype(np.int64)
        order = np.argsort(d, kind="stable")
        raw = np.ascontiguousarray(recs[order]).view(np.uint8).reshape(-1)
        send[: raw.size] = torch.from_numpy(raw.copy())
        return np.bincount(d, minlength=world).astype(np.int64)

    def agg_replace_device(self, n, recv, stream=0):
        arr = recv[: n * 48].numpy().view(JREC_DTYPE).copy()
        self.recs = [arr[i] for i in range(n)]

    def agg_finalize(self, stream=0):
        acc = {}
        for r in sorted(self.recs, key=lambda x: int(x["idx"])):  # stream order, whatever the arrival order was
            key = (int(r["chrom"]), int(r["start"]), int(r["end"]), int(r["sk"]) & 3)
            a = acc.setdefault(key, dict(first=int(r["idx"]), w=0.0, b=0.0, n=0, ql=[], qr=[], d=[], o=[], nh=[], rh=set(),
                                         qh=set(), pal=set(), sig=0))
            wt = 1.0 / ((int(r["sk"]) >> 8) & 0xFF)
            a["w"] += wt
            if r["q_left"] != 0 and r["q_right"] != 0:
                a["b"] += wt
            a["n"] += 1
            a["ql"].append(int(r["q_left"])); a["qr"].append(int(r["q_right"]))
            a["d"].append(int(r["dist"])); a["o"].append(int(r["ov"])); a["nh"].append(int(r["n_hits"]))
            a["rh"].add(int(r["read_hash"]))
            if int(r["read_hash"]) & 1:
                a["pal"].add(int(r["read_hash"]))
            a["qh"].add(int(r["qname_hash"]))
            a["sig"] = (int(r["sk"]) >> 16) & 0xFFF
        out = np.zeros(len(acc), dtype=JUNCTION_DTYPE)
        for k, (key, a) in enumerate(sorted(acc.items(), key=lambda kv: kv[1]["first"])):
            out[k] = (key[0], key[1], key[2], key[3] | (a["sig"] << 16), a["first"], a["w"], a["b"], a["n"], len(a["qh"]),
                      len(a["rh"]) - (len(a["pal"]) + 1) // 2, max(a["ql"]), max(a["qr"]), min(a["nh"]), min(a["d"]), min(a["o"]), 0)
        self.junc = out
        return len(out)

    def agg_fetch(self, n):
        return self.junc[:n]

    def close(self):
        pass
