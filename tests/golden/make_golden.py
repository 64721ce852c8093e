#!/usr/bin/env python3
"""
Generate the golden fixtures under tests/golden/ by running the REFERENCE's own find_circ.py
(through oracle/ref_shim/run_reference.py -- see that file for the py2->py3 mechanics) on inputs
built here.  Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

Inputs (genome FASTA + SAM) and the reference's outputs are committed; the tests never need
/root/reference.  Cases:

  kat3        reference test_data/test_ref.fa + test_reads.fa, segments placed by exact matching
              (stands in for `bwa mem -k 15 -T 1`, test_data/Makefile:14-15), shipped defaults
  cdr1as      reference test_data/CDR1as_locus.fa + cdr1as_reads.fa as BWA-style 2-segment reads
  synth_*     seeded synthetic genome with planted back-splice / linear junctions, 2-segment reads with
              errors, N's, chromosome-edge windows, non-unique anchors, plus hand-assembled multi-segment
              and paired-end fragments; run under several option sets
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from find_circ2_b200 import synth  # noqa: E402

REF_DATA = "/root/reference/test_data"
SHIM = os.path.join(ROOT, "oracle", "ref_shim", "run_reference.py")


# ----------------------------------------------------------------------------------------------
def read_fasta(path):
    names, seqs = [], []
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            names.append(line[1:].split()[0])
            seqs.append([])
        elif line:
            seqs[-1].append(line)
    return names, ["".join(s) for s in seqs]


def exact_segments(read, chroms, min_len=12):
    """Greedy longest-prefix exact matching of `read` (genome orientation tried both ways) -- a stand-in
    for the aligner on the reference's tiny test sets.  Returns (is_reverse, [Seg...]) or None."""
    best = None
    for rev in (False, True):
        r = synth.revcomp(read) if rev else read
        segs, q = [], 0
        ok = True
        while q < len(r):
            hit = None
            for ln in range(len(r) - q, min_len - 1, -1):
                sub = r[q : q + ln]
                for cname, cseq in chroms:
                    p = cseq.find(sub)
                    if p >= 0:
                        hit = (cname, p, ln)
                        break
                if hit:
                    break
            if not hit:
                ok = False
                break
            segs.append(synth.Seg(hit[0], hit[1], q, q + hit[2], reverse=rev))
            q += hit[2]
        if ok and (best is None or len(segs) < len(best[1])):
            best = (rev, segs, r)
    return best


def build_ref_case(genome_fa, reads_fa, out_dir, name):
    os.makedirs(out_dir, exist_ok=True)
    shutil.copy(genome_fa, os.path.join(out_dir, "genome.fa"))
    gnames, gseqs = read_fasta(genome_fa)
    chroms = [(n, s.upper()) for n, s in zip(gnames, gseqs)]
    rnames, rseqs = read_fasta(reads_fa)
    with open(os.path.join(out_dir, "input.sam"), "w") as fh:
        fh.write("@HD\tVN:1.3\tSO:unsorted\n")
        for n, s in zip(gnames, gseqs):
            fh.write("@SQ\tSN:%s\tLN:%d\n" % (n, len(s)))
        for qn, rs in zip(rnames, rseqs):
            m = exact_segments(rs.upper(), chroms)
            if m is None:
                fh.writelines(synth.records_for_read(qn, rs, [], unmapped=True))
                continue
            rev, segs, oriented = m
            # primary = longest segment, listed first like bwa does
            order = sorted(range(len(segs)), key=lambda k: -(segs[k].q1 - segs[k].q0))
            segs = [segs[k] for k in order]
            fh.writelines(synth.records_for_read(qn, oriented, segs))


# ----------------------------------------------------------------------------------------------
def build_synth_case(out_dir, seed=11, n_pairs=900, read_len=100, asize=15, error_rate=0.01, paired_extra=True):
    os.makedirs(out_dir, exist_ok=True)
    g = synth.make_genome([30000, 22000, 9000], seed=seed, n_frac=0.01, n_run=(20, 200), soft_frac=0.05, soft_run=(50, 400))
    J = synth.plant_junctions(g, n_circ=50, n_lin=30, seed=seed + 1, span=(150, 6000), margin=300)
    t = synth.make_pairs(
        g, J, n_pairs, read_len=read_len, asize=asize, seed=seed + 2, error_rate=error_rate,
        frac_decoy=0.12, frac_nonuniq=0.06, frac_edge=0.04, frac_inner_shift=0.2, frac_read_n=0.03, frac_no_xs=0.1,
    )
    rng = np.random.default_rng(seed + 3)
    lines = []  # list of per-fragment record lists, shuffled at the end (fragments stay contiguous)

    for i in range(len(t)):
        qual = "I" * read_len if (i % 3 == 0) else None
        lines.append(synth.bwa_records_for_pair(g, t, i, "r%d" % i, qual))
        # PCR duplicate (same sequence, new name) and same-name re-occurrence exercise n_uniq / n_frags
        if i % 37 == 0:
            lines.append(synth.bwa_records_for_pair(g, t, i, "dup%d" % i, qual))

    if paired_extra:
        lines += complex_fragments(g, J, rng, read_len, asize)

    g.write_fasta(os.path.join(out_dir, "genome.fa"))  # after the last planting
    order = rng.permutation(len(lines))
    with open(os.path.join(out_dir, "input.sam"), "w") as fh:
        fh.write(synth.sam_header(g))
        # an unmapped record first: the reference never checks the first record (find_circ.py:1462-1463)
        fh.writelines(synth.records_for_read("first_unmapped", "ACGT" * 10, [], unmapped=True))
        for k in order:
            fh.writelines(lines[k])
            if k % 50 == 0:
                fh.writelines(synth.records_for_read("um%d" % k, "ACGTN" * 10, [], unmapped=True))
    return g, J, t


def complex_fragments(g, J, rng, R, asize):
    """Multi-segment and paired-end fragments that reach every branch of record_hits (find_circ.py:1276-1439)."""
    out = []
    gs = lambda c, s, e: synth.genome_slice(g, g.names[c], s, e)  # noqa: E731
    circs = [k for k in range(len(J)) if J.circ[k]]
    lins = [k for k in range(len(J)) if not J.circ[k]]
    n = 0

    def nm(tag):
        nonlocal n
        n += 1
        return "cx_%s_%d" % (tag, n)

    for rep in range(6):
        # (1) closure: a read that runs around a short circle: tail | whole circle | head
        k = circs[rep]
        c, s, e = int(J.chrom[k]), int(J.start[k]), int(J.end[k])
        cname = g.names[c]
        span = e - s
        body = min(span, 60)
        # synthetic short circle made of the last `body` bases before e ... we need the true circle to be short,
        # so use a dedicated window: take the circle [e-body, e) only if the left signal exists -> otherwise it is a no-bp case
        a = 20 + rep
        read = gs(c, e - a, e) + gs(c, s, s + min(span, R - 2 * a)) + gs(c, s, s + a)
        if span <= R - 2 * a:
            mid = span
            read = gs(c, e - a, e) + gs(c, s, e) + gs(c, s, s + a)
            segs = [synth.Seg(cname, s, a, a + mid), synth.Seg(cname, e - a, 0, a), synth.Seg(cname, s, a + mid, a + mid + a)]
            out.append(synth.records_for_read(nm("closure"), read, segs))

        # (2) paired-end: mate1 across the back-splice, mate2 unspliced inside / outside / other chromosome
        j = 30 + 5 * rep
        m1 = gs(c, e - j, e) + gs(c, s, s + R - j)
        segs1 = [synth.Seg(cname, e - j, 0, j, reverse=bool(rep & 1)), synth.Seg(cname, s, j, R, reverse=bool(rep & 1))]
        where = rep % 3
        if where == 0:
            p2 = s + max(0, (span - 60) // 2)
            c2 = c
        elif where == 1:
            p2 = e + 100
            c2 = c
        else:
            c2 = (c + 1) % len(g.names)
            p2 = 1000 + 10 * rep
        m2 = gs(c2, p2, p2 + 60)
        q = nm("pe")
        recs = synth.records_for_read(q, m1, segs1, mate=1) + synth.records_for_read(
            q, m2, [synth.Seg(g.names[c2], p2, 0, 60, reverse=not bool(rep & 1))], mate=2
        )
        out.append(recs)

        # (3) both mates cross the same back-splice -> two circ spans with the same coordinate (SUPPORT_CLOSURE)
        j2 = R - j
        m2b = gs(c, e - j2, e) + gs(c, s, s + R - j2)
        q = nm("pe2")
        recs = synth.records_for_read(q, m1, segs1, mate=1) + synth.records_for_read(
            q, m2b, [synth.Seg(cname, e - j2, 0, j2), synth.Seg(cname, s, j2, R)], mate=2
        )
        out.append(recs)

        # (4) two different back-splices in one fragment -> WARN_MULTI_BACKSPLICE
        k2 = circs[rep + 7]
        if int(J.chrom[k2]) == c or True:
            c_b, s_b, e_b = int(J.chrom[k2]), int(J.start[k2]), int(J.end[k2])
            m2c = gs(c_b, e_b - 40, e_b) + gs(c_b, s_b, s_b + R - 40)
            q = nm("multi")
            recs = synth.records_for_read(q, m1, segs1, mate=1) + synth.records_for_read(
                q, m2c, [synth.Seg(g.names[c_b], e_b - 40, 0, 40), synth.Seg(g.names[c_b], s_b, 40, R)], mate=2
            )
            out.append(recs)

        # (5) linear junction in mate2 inside / outside the circle of mate1
        kl = lins[rep]
        cl, sl, el = int(J.chrom[kl]), int(J.start[kl]), int(J.end[kl])
        ml = gs(cl, sl - 45, sl) + gs(cl, el, el + R - 45)
        q = nm("lin_with_circ")
        # force the linear read onto the circ's chromosome coordinates only when they share a chromosome
        recs = synth.records_for_read(q, m1, segs1, mate=1) + synth.records_for_read(
            q, ml, [synth.Seg(g.names[cl], sl - 45, 0, 45), synth.Seg(g.names[cl], el, 45, R)], mate=2
        )
        out.append(recs)

        # (6) three linear segments in one read (two introns) and a read that is linear then back-spliced
        if rep + 1 < len(lins):
            kl2 = lins[rep + 1]
            cl2, sl2, el2 = int(J.chrom[kl2]), int(J.start[kl2]), int(J.end[kl2])
            read3 = gs(cl, sl - 30, sl) + gs(cl, el, el + 40) + gs(c, s, s + 30)
            segs3 = [
                synth.Seg(g.names[cl], el, 30, 70),
                synth.Seg(g.names[cl], sl - 30, 0, 30),
                synth.Seg(cname, s, 70, 100, reverse=False),
            ]
            out.append(synth.records_for_read(nm("threeseg"), read3, segs3))

        # (7) broken read: two proper segments cover only the first 60 nt, the rest maps to another strand / chromosome
        j3 = 25 + rep
        readb = gs(c, e - j3, e) + gs(c, s, s + 60 - j3) + synth.revcomp(gs(c, s + 200, s + 240))
        segsb = [
            synth.Seg(cname, e - j3, 0, j3),
            synth.Seg(cname, s, j3, 60),
            synth.Seg(cname, s + 200, 60, 100, reverse=True),
        ]
        out.append(synth.records_for_read(nm("broken_strand"), readb, segsb))
        c3 = (c + 2) % len(g.names)
        readc = gs(c, e - j3, e) + gs(c, s, s + 60 - j3) + gs(c3, 500, 540)
        segsc = [synth.Seg(cname, e - j3, 0, j3), synth.Seg(cname, s, j3, 60), synth.Seg(g.names[c3], 500, 60, 100)]
        out.append(synth.records_for_read(nm("broken_chrom"), readc, segsc))

        # (8) a segment shorter than the anchor size
        short = max(asize - 3, 5)
        reads_ = gs(c, e - short, e) + gs(c, s, s + R - short)
        out.append(
            synth.records_for_read(nm("short"), reads_, [synth.Seg(cname, s, short, R), synth.Seg(cname, e - short, 0, short)])
        )

        # (9) overlapping segments as bwa reports them (56M20S + 52H24M, test_data/test_norm.sam:94-95)
        jo = 50
        ro = gs(c, e - jo, e) + gs(c, s, s + R - jo)
        out.append(
            synth.records_for_read(nm("overlap"), ro, [synth.Seg(cname, e - jo, 0, jo + 3), synth.Seg(cname, s - 4, jo - 4, R)])
        )

        # (10) ambiguous breakpoint: tandem GTAG homology gives two canonical splits with equal score
        # handled by the dedicated generator below
    out += ambiguous_fragments(g, rng, R)
    return out


def ambiguous_fragments(g, rng, R):
    """Plant 'GTAG' tandem homology so that two split positions 4 nt apart both carry GT/AG and explain the read
    equally well -> n_hits 2, WARN_AMBIGUOUS_BP (find_circ.py:966-974, 613-614)."""
    out = []
    c = 0
    seq = g.seqs[c]
    for rep in range(4):
        s = 3000 + rep * 2500
        e = s + 800 + rep * 50
        # donor side: exon | GTAGGT ; acceptor side: AG | GTAG exon   (0-based: seq[e:e+6], seq[s-2:s+4]):
        # the split x (true) and x+4 both show GT/AG and the 4 read bases in between match either flank
        seq[e : e + 6] = np.frombuffer(b"GTAGGT", dtype=np.uint8)
        seq[s - 2 : s + 4] = np.frombuffer(b"AGGTAG", dtype=np.uint8)
        j = 40 + rep
        read = synth.genome_slice(g, g.names[c], e - j, e) + synth.genome_slice(g, g.names[c], s, s + R - j)
        segs = [synth.Seg(g.names[c], e - j, 0, j), synth.Seg(g.names[c], s, j, R)]
        out.append(synth.records_for_read("ambig_%d" % rep, read, segs))
        # second read of the same junction, shifted
        j = 55 + rep
        read = synth.genome_slice(g, g.names[c], e - j, e) + synth.genome_slice(g, g.names[c], s, s + R - j)
        segs = [synth.Seg(g.names[c], s, j, R), synth.Seg(g.names[c], e - j, 0, j)]
        out.append(synth.records_for_read("ambig_b_%d" % rep, read, segs))
    return out


# ----------------------------------------------------------------------------------------------
def run_reference(case_dir, tag, args):
    """Run the shimmed reference on case_dir/input.sam; store outputs under case_dir/ref_<tag>/."""
    out = os.path.join(case_dir, "ref_" + tag)
    if os.path.isdir(out):
        shutil.rmtree(out)
    with tempfile.TemporaryDirectory() as tmp:
        # the reference writes <genome>.byo_index next to the FASTA (find_circ.py:110-115)
        gfa = os.path.join(tmp, "genome.fa")
        shutil.copy(os.path.join(case_dir, "genome.fa"), gfa)
        run_dir = os.path.join(tmp, "run")
        real = [a.replace("@CASE@", case_dir) for a in args]  # cmdline.txt keeps the placeholder, the tests expand it
        cmd = [sys.executable, SHIM, "-G", gfa, "-o", run_dir, "-q"] + real + [os.path.join(case_dir, "input.sam")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        os.makedirs(out)
        with open(os.path.join(out, "cmdline.txt"), "w") as fh:
            fh.write(" ".join(args) + "\n")
            fh.write("exit=%d\n" % r.returncode)
        if r.returncode != 0:
            with open(os.path.join(out, "stderr.txt"), "w") as fh:
                fh.write(r.stderr[-4000:])
            print("  [%s] exit %d" % (tag, r.returncode))
            return
        for fn in ("circ_splice_sites.bed", "lin_splice_sites.bed", "multi_events.tsv"):
            shutil.copy(os.path.join(run_dir, fn), os.path.join(out, fn))
        if os.path.exists(os.path.join(run_dir, "test_results.tsv")):  # --test
            shutil.copy(os.path.join(run_dir, "test_results.tsv"), os.path.join(out, "test_results.tsv"))
        if r.stdout:  # --stdout redirects one of the outputs (find_circ.py:453-458)
            with open(os.path.join(out, "stdout.txt"), "w") as fh:
                fh.write(r.stdout)
        with gzip.open(os.path.join(run_dir, "spliced_reads.fastq.gz"), "rt") as fin, open(
            os.path.join(out, "spliced_reads.fastq"), "w"
        ) as fo:
            try:
                fo.write(fin.read())
            except EOFError:
                # --stdout reads: the reference writes one line into the GzipFile and never closes it (find_circ.py:456-458)
                fo.write("# (truncated gzip stream in the reference)\n")
        # keep only the counters of run.log (timestamps are not reproducible)
        with open(os.path.join(out, "counters.txt"), "w") as fo:
            seen = False
            for line in open(os.path.join(run_dir, "run.log")):
                msg = line.rstrip("\n").split("\t")[-1]
                if msg == "run finished":
                    seen = True
                    continue
                if seen and "=" in msg:
                    fo.write(msg + "\n")
        nc = sum(1 for _ in open(os.path.join(out, "circ_splice_sites.bed"))) - 1
        nl = sum(1 for _ in open(os.path.join(out, "lin_splice_sites.bed"))) - 1
        print("  [%s] circ=%d lin=%d" % (tag, nc, nl))


OPTION_SETS = {
    "default": ["-n", "test"],
    "a20": ["-n", "smp", "-a", "20"],
    "noncanonical": ["-n", "test", "--non-canonical"],
    "allhits": ["-n", "test", "--all-hits"],
    "noncanonical_allhits": ["-n", "test", "--non-canonical", "--all-hits"],
    "strandpref": ["-n", "test", "--strand-pref", "--non-canonical"],
    "m0_d1": ["-n", "test", "-m", "0", "-d", "1"],
    "m4_d3": ["-n", "test", "-m", "4", "-d", "3"],
    "d0": ["-n", "test", "-d", "0"],
    "uniq0_half_nobridge": ["-n", "test", "--min-uniq-qual", "0", "--half-unique", "--report-nobridges"],
    "uniq0": ["-n", "test", "--min-uniq-qual", "0"],
    "nolinear_nomulti": ["-n", "test", "--no-linear", "--no-multi"],
    "thresholds": ["-n", "test", "--short-threshold", "400", "--huge-threshold", "3000"],
    "known": ["-n", "test", "--known-circ", "@CASE@/known_circ.bed", "--known-lin", "@CASE@/known_lin.bed"],
}


def write_known_sites(case_dir):
    """BED6 files for --known-circ / --known-lin (find_circ.py:665-679) from the default run of the case: every third
    junction under its own name, one with the strand flipped and a few sites that no read supports"""
    for kind, fn in (("circ", "circ_splice_sites.bed"), ("lin", "lin_splice_sites.bed")):
        rows = [l.split("\t") for l in open(os.path.join(case_dir, "ref_default", fn)) if not l.startswith("#")]
        with open(os.path.join(case_dir, "known_%s.bed" % kind), "w") as fh:
            fh.write("# known %s junctions\n" % kind)
            for k, r in enumerate(rows):
                if k % 3 == 0:
                    fh.write("\t".join([r[0], r[1], r[2], "KNOWN_%s_%d" % (kind.upper(), k), "0", r[5]]) + "\n")
                elif k % 7 == 1:
                    fh.write("\t".join([r[0], r[1], r[2], "FLIPPED_%d" % k, "0", "-" if r[5] == "+" else "+"]) + "\n")
                elif k % 7 == 2:
                    fh.write("\t".join([r[0], str(int(r[1]) + 1), r[2], "SHIFTED_%d" % k, "0", r[5]]) + "\n")
            fh.write("chrNotThere\t10\t500\tELSEWHERE\t0\t+\n")


def main():
    if sys.argv[1:] == ["cli"]:  # rarely used switches of the command line (added after the first set)
        d = os.path.join(HERE, "synth_b")
        run_reference(d, "noop", ["-n", "test", "--noop"])
        run_reference(d, "stdout_circs", ["-n", "test", "--stdout", "circs"])
        run_reference(d, "stdout_reads", ["-n", "test", "--stdout", "reads", "-t", "--chunk-size", "100"])
        return
    if sys.argv[1:] == ["selftest"]:  # --test: the reference's own validation against the truth in the read names
        run_reference(os.path.join(HERE, "kat3"), "selftest", ["-n", "test", "--test"])
        run_reference(os.path.join(HERE, "kat3"), "selftest_a20", ["-n", "test", "-a", "20", "--test"])
        run_reference(os.path.join(HERE, "synth_a"), "selftest", ["-n", "test", "--test"])
        return
    if sys.argv[1:] == ["known"]:  # only the runs with known junctions (added after the first set)
        d = os.path.join(HERE, "synth_a")
        write_known_sites(d)
        run_reference(d, "known", OPTION_SETS["known"])
        run_reference(d, "known_d0", OPTION_SETS["known"] + ["-d", "0", "--min-uniq-qual", "0", "--report-nobridges"])
        return
    print("kat3")
    d = os.path.join(HERE, "kat3")
    build_ref_case(os.path.join(REF_DATA, "test_ref.fa"), os.path.join(REF_DATA, "test_reads.fa"), d, "kat3")
    run_reference(d, "default", ["-n", "test"])
    run_reference(d, "a20", ["-n", "test", "-a", "20"])

    print("cdr1as")
    d = os.path.join(HERE, "cdr1as")
    build_ref_case(os.path.join(REF_DATA, "CDR1as_locus.fa"), os.path.join(REF_DATA, "cdr1as_reads.fa"), d, "cdr1as")
    run_reference(d, "default", ["-n", "test"])
    run_reference(d, "a20", ["-n", "test", "-a", "20"])

    print("synth_a")
    d = os.path.join(HERE, "synth_a")
    build_synth_case(d, seed=11, n_pairs=900, read_len=100, asize=15, error_rate=0.01)
    for tag, args in OPTION_SETS.items():
        if tag == "known":
            write_known_sites(d)
        run_reference(d, tag, args)
    run_reference(d, "known_d0", OPTION_SETS["known"] + ["-d", "0", "--min-uniq-qual", "0", "--report-nobridges"])

    print("synth_b (150-nt reads, 2% errors)")
    d = os.path.join(HERE, "synth_b")
    build_synth_case(d, seed=23, n_pairs=500, read_len=150, asize=20, error_rate=0.02, paired_extra=False)
    for tag in ("default", "a20", "noncanonical_allhits", "uniq0_half_nobridge"):
        run_reference(d, tag, OPTION_SETS[tag])


if __name__ == "__main__":
    main()
