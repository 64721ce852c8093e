"""
merge_bed.py.  tests/golden/merge/ref_*.out are outputs of the REFERENCE's merge_bed.py
(oracle/ref_shim/run_merge_bed.py, tests/golden/make_golden_merge.py) in every mode.  CPU: the oracle restatement
(oracle/merge_bed_oracle.py) is pinned to them; GPU (-m gpu): the product (find_circ2_b200/merge_bed.py, keyed merge by
fc_merge_tables on the device) reproduces them.  Rows are compared after a sort (python-2 dict order upstream).
"""
import os

import pytest

from conftest import GOLDEN
from oracle import merge_bed_oracle as MO

MERGE = os.path.join(GOLDEN, "merge")
RUNS = sorted(f[4:-4] for f in os.listdir(MERGE) if f.startswith("ref_") and f.endswith(".cmd"))


def _args(tag):
    argv = open(os.path.join(MERGE, "ref_%s.cmd" % tag)).read().split()
    flags = dict(bed6="-6" in argv, score="--score" in argv, verbatim="-V" in argv)
    return [os.path.join(MERGE, a) for a in argv if not a.startswith("-")], flags


@pytest.mark.parametrize("tag", RUNS)
def test_oracle_reproduces_reference_merge(tag):
    paths, flags = _args(tag)
    text, stats = MO.merge(paths, **flags)
    want = open(os.path.join(MERGE, "ref_%s.out" % tag)).read()
    assert MO.canonical(text, flags["score"]) == MO.canonical(want, flags["score"])
    assert stats == open(os.path.join(MERGE, "ref_%s.stats" % tag)).read()
    assert len(want.splitlines()) > 100


@pytest.mark.gpu
@pytest.mark.parametrize("tag", RUNS)
def test_gpu_merge_reproduces_reference_merge(tag):
    from find_circ2_b200 import merge_bed

    paths, flags = _args(tag)
    text, stats = merge_bed.merge_to_text(paths, **flags)
    want = open(os.path.join(MERGE, "ref_%s.out" % tag)).read()
    assert MO.canonical(text, flags["score"]) == MO.canonical(want, flags["score"])
    assert stats == open(os.path.join(MERGE, "ref_%s.stats" % tag)).read()
