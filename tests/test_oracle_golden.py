"""
Pins the CPU oracle (oracle/find_circ_oracle.py) to the reference: every output file of every
reference run under tests/golden/ (produced by /root/reference/find_circ.py itself, see
tests/golden/make_golden.py) must be reproduced byte for byte after the canonical row sort.
"""
import os

import pytest

from conftest import golden_cases, golden_ids
import helpers as H
from oracle import find_circ_oracle as O


@pytest.mark.parametrize("case_dir,ref_dir,argv", golden_cases(), ids=golden_ids())
def test_oracle_matches_reference(case_dir, ref_dir, argv):
    opt = O.options_from_argv(argv)
    out = O.run(os.path.join(case_dir, "genome.fa"), os.path.join(case_dir, "input.sam"), opt)
    H.compare_outputs(out.circ_bed, out.lin_bed, out.reads_fastq, out.multi_events, out.counters, ref_dir, argv, out.test_results)


def test_kat1_scan_level():
    """the CDR1as reads at a=20,m=2,d=2 (test_data/cdr1as_reference.bed:2 -> 728..2213 +)"""
    g = O.Genome(os.path.join(os.path.dirname(__file__), "golden", "cdr1as", "genome.fa"))
    reads = {}
    name = None
    for line in open(os.path.join(os.path.dirname(__file__), "golden", "cdr1as", "reads.fa")):
        if line.startswith(">"):
            name = line[1:].strip()
        elif line.strip():
            reads[name] = line.strip()
    opt = O.Options(asize=20)
    expect = {"r1": (2161, 752, 34), "r2": (2191, 782, 4), "r4": (2193, 784, 2), "r3": (2152, 743, None)}
    for rn, (a_pos, b_aend, x) in expect.items():
        read = reads[rn]
        a_start, b_end, l = O.window_geometry(a_pos, b_aend, len(read), opt)
        hits = O.scan_windows(
            g.get("CDR1as_locus", a_start, a_start + l + 2).upper(),
            g.get("CDR1as_locus", b_end - l - 2, b_end).upper(),
            read[18:-18], "CDR1as_locus", a_start, b_end, True, "+", opt,
        )
        if x is None:
            assert hits == []
        else:
            assert len(hits) == 1
            h = hits[0]
            assert (h.start, h.end, h.strand, int(h.dist), h.ov, h.gtag, h.n_hits) == (728, 2213, "+", 0, 0, "GTAG", 1)
