"""the v1.2 face (bowtie2 anchor pairs in, 20-column BED out) against the reference's one golden line
(test_data/cdr1as_reference.bed:2) -- host logic on CPU with the test double, the real engine under -m gpu"""
import os

import pytest

from conftest import GOLDEN


def _run(engine=None):
    from find_circ2_b200 import cli

    d = os.path.join(GOLDEN, "cdr1as")
    argv = ["-G", os.path.join(d, "genome.fa"), "-n", "test", "-p", "cdr1as_test_", "--stats", "/dev/null", "--reads", "/dev/null"]
    assert cli.wants_v12(argv)
    opt = cli.parse_args_v12(argv)[0]
    if engine is not None:
        engine = engine(0, opt.asize, opt.margin, opt.maxdist, False, False)
        engine.load_genome_fasta(opt.genome)
    return cli.run_v12_to_strings(opt, os.path.join(d, "anchors.sam"), engine=engine)


def _check(out):
    golden = open(os.path.join(GOLDEN, "cdr1as", "cdr1as_reference.bed")).read().split("\n")
    got = out["bed"].split("\n")
    assert got[0] == golden[0]           # header of the circular section
    assert got[1] == golden[1]           # THE golden row
    assert got[2] == golden[2]           # header of the (empty) linear section
    assert out["reads"].count(">") == 3  # r1, r2, r4 support the junction; r3's breakpoint is 5 nt inside anchor B
    assert "circ_reads\t3" in out["stats"] and "circ_no_bp\t1" in out["stats"]


def test_v12_golden_line_host_logic():
    from fake_engine import FakeEngine

    _check(_run(FakeEngine))


@pytest.mark.gpu
def test_v12_golden_line_gpu():
    _check(_run())


def test_cli_dispatch():
    from find_circ2_b200 import cli

    assert not cli.wants_v12(["-G", "g.fa", "-q", "-n", "x", "in.sam"])
    assert cli.wants_v12(["-G", "g.fa", "-p", "pre_", "-q", "3"])
    o = cli.parse_args_v12(["-G", "g.fa", "-p", "pre_", "-q", "3", "--halfuniq", "--report_nobridge"])[0]
    assert (o.min_uniq_qual, o.halfunique, o.report_nobridges, o.asize) == (3, True, True, 20)
