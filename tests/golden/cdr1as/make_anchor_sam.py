#!/usr/bin/env python3
"""bowtie2-style anchor records for the reference's CDR1as reads (KAT-1/KAT-2 of SURVEY.md 8c): the 20-nt anchors of
test_data/cdr1as_reads.fa named as unmapped2anchors.py:124-132 does, placed where they align on CDR1as_locus.fa
(r2's A anchor and r3's B anchor carry one mismatch).  Writes anchors.sam next to this script."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
reads = {}
name = None
for line in open(os.path.join(HERE, "reads.fa")):
    if line.startswith(">"):
        name = line[1:].strip()
    elif line.strip():
        reads[name] = line.strip()
# (A.pos, B.aend, mismatches in A, mismatches in B), 0-based, from SURVEY.md 8c KAT-2
place = {"r1": (2161, 752, 0, 0), "r2": (2191, 782, 1, 0), "r3": (2152, 743, 0, 1), "r4": (2193, 784, 0, 0)}
with open(os.path.join(HERE, "anchors.sam"), "w") as fh:
    fh.write("@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:CDR1as_locus\tLN:2924\n")
    for rn in ("r1", "r2", "r3", "r4"):
        seq = reads[rn]
        a_pos, b_aend, mma, mmb = place[rn]
        for tag, pos, s, mm in (("%s_A__%s" % (rn, seq), a_pos, seq[:20], mma), ("%s_B" % rn, b_aend - 20, seq[-20:], mmb)):
            fh.write("%s\t0\tCDR1as_locus\t%d\t42\t20M\t*\t0\t0\t%s\t%s\tAS:i:%d\tXN:i:0\tNM:i:%d\n" % (tag, pos + 1, s, "I" * 20, -6 * mm, mm))
