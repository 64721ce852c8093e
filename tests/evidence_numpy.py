"""
TEST TWIN of fc_ingest_evidence (csrc/ingest.cu): the evidence rules of record_hits (find_circ.py:1276-1439) for a batch of
fragments with at most two spans, written with numpy column operations (this was the product's implementation before the
rules moved to C++; kept as an independent statement of the same rules).  Never imported by the product.
"""
import numpy as np

FR_UNSPLICED, FR_OTHER_CHROM, FR_BROKEN, FR_TWO_MATES = 1, 2, 4, 8


def evidence_numpy(a, hits, n, m, off, asize, B):
    has = (hits["w2"] & 0xFFFF) > 0
    state, kind, ff = a["f_state"][:m], a["f_kind"][:m], a["f_flags"][:m]
    row0 = a["f_row0"][:m].astype(np.int64)
    two = a["f_nsp"][:m] == 2
    queued = [(state & 1) > 0, (state & 2) > 0]
    row = [np.where(queued[0], row0, 0), np.where(queued[1], row0 + (state & 1), 0)]
    circ = [(kind & 1) > 0, (kind & 2) > 0]
    lin = [~circ[0], two & ~circ[1]]
    hit = [queued[j] & has[row[j]] for j in (0, 1)]
    count = np.count_nonzero
    counters = [sum(count(circ[j] & hit[j]) for j in (0, 1)), sum(count(circ[j] & queued[j] & ~hit[j]) for j in (0, 1)),
                sum(count(lin[j] & hit[j]) for j in (0, 1)), sum(count(lin[j] & queued[j] & ~hit[j]) for j in (0, 1))]
    chrom = a["chrom"][:n].astype(np.int64)
    h_start, h_end, h_minus = hits["start"].astype(np.int64), hits["end"].astype(np.int64), (hits["w3"] & 1).astype(np.int64)
    key = [np.stack([chrom[row[j]], h_start[row[j]], h_end[row[j]], h_minus[row[j]], lin[j].astype(np.int64)], axis=1) for j in (0, 1)]
    ch = [circ[j] & hit[j] for j in (0, 1)]
    both = ch[0] & ch[1]
    multi = both & (key[0] != key[1]).any(axis=1)
    circ_any = ch[0] | ch[1]
    single = circ_any & ~multi
    ck = np.where(ch[1][:, None], key[1], key[0])
    cs, ce = ck[:, 1], ck[:, 2]
    W = np.zeros(m, dtype=np.uint32)

    def flag(cond, name):
        W[cond] |= np.uint32(B[name])

    flag((circ[0] & queued[0] & ~hit[0]) | (circ[1] & queued[1] & ~hit[1]), "WARN_UNRESOLVED_EXTRA_BACKSPLICE")
    flag(single & circ[0] & circ[1], "SUPPORT_CLOSURE")
    lin_ev, lin_out = [], []
    for j in (0, 1):
        flag(lin[j] & queued[j] & ~hit[j], "WARN_UNRESOLVED_LINSPLICE")
        ev = lin[j] & hit[j] & single
        outside = (key[j][:, 1] <= cs) | (key[j][:, 2] >= ce)
        flag(ev & outside, "WARN_OUTSIDE_SPLICE_JUNCTION")
        flag(ev & ~outside, "SUPPORT_INSIDE_SPLICE_JUNCTION")
        lin_ev.append(ev)
        lin_out.append(ev & outside)
    un = ((ff & FR_UNSPLICED) > 0) & single
    un_other = un & ((ff & FR_OTHER_CHROM) > 0)
    un_pos, un_aend = a["f_un_pos"][:m].astype(np.int64), a["f_un_aend"][:m].astype(np.int64)
    un_outside = un & ~un_other & ((un_pos + asize <= cs) | (un_aend - asize >= ce))
    flag(un_other, "WARN_OTHER_CHROM_MATE")
    flag(un_outside, "WARN_OUTSIDE_MATE")
    flag(un & ~un_other & ~un_outside, "SUPPORT_INSIDE_MATE")
    flag(single & ((ff & FR_BROKEN) > 0), "BROKEN_SEGMENTS")
    W[multi] = B["WARN_MULTI_BACKSPLICE"]
    name_hash = a["qname_hash"][:n][row0]
    ev1 = np.nonzero(single & (W != 0))[0]
    evm = np.nonzero(multi)[0]
    ev_key = np.concatenate([ck[ev1], key[0][evm], key[1][evm]])
    ev_hash = np.concatenate([name_hash[ev1], name_hash[evm], name_hash[evm]])
    ev_mask = np.concatenate([W[ev1], W[evm], W[evm]])
    cls = (hit[0] * 1 + hit[1] * 2 + lin_ev[0] * 4 + lin_ev[1] * 8 + lin_out[0] * 16 + lin_out[1] * 32 + un * 64
           + (un_other | un_outside) * 128).astype(np.uint8)
    txt_off = a["f_txt_off"][:6 * m].reshape(m, 2, 3)
    txt_len = a["f_txt_len"][:6 * m].reshape(m, 2, 3)
    fr = np.nonzero(hit[0] | hit[1])[0]
    first = np.where(hit[0][fr, None], key[0][fr], key[1][fr])
    second = np.where((hit[0] & hit[1])[fr, None] & (key[0][fr] != key[1][fr]).any(axis=1)[:, None], key[1][fr], -1)
    n_mates = 1 + ((ff[fr] & FR_TWO_MATES) > 0)
    who = np.repeat(np.arange(len(fr)), n_mates)
    mate = np.arange(len(who)) - np.repeat(np.cumsum(n_mates) - n_mates, n_mates)
    return dict(counters=[int(c) for c in counters], any_hit=bool(hit[0].any() or hit[1].any()), W=W, cls=cls, key0=key[0], key1=key[1], ck=ck,
                ev_key=ev_key, ev_hash=ev_hash, ev_mask=ev_mask, r_seq=a["f_seq"][:m][fr[who]], r_k0=first[who], r_k1=second[who],
                r_mask=W[fr[who]].astype(np.int64), r_off3=txt_off[fr[who], mate] + off, r_len3=txt_len[fr[who], mate])
