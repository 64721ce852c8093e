"""
End-to-end drop-in parity on the GPU: the product pipeline (host logic + CUDA scan + CUDA aggregation) must
reproduce every output file the REFERENCE produced (tests/golden/*/ref_*) byte for byte after the canonical sort,
for every stored option set.
"""
import os

import pytest

from conftest import golden_cases, golden_ids
from oracle import find_circ_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case_dir,ref_dir,argv", golden_cases(), ids=golden_ids())
@pytest.mark.parametrize("batch,native", [(1 << 18, True), (97, False), (1500, True)])
def test_pipeline_matches_reference(case_dir, ref_dir, argv, batch, native):
    from find_circ2_b200 import cli

    opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv + [os.path.join(case_dir, "input.sam")])[0]
    opt.batch_pairs = batch
    if native and opt.allhits:
        native = False  # --all-hits always uses the python ingest
    out = cli.run_to_strings(opt, os.path.join(case_dir, "input.sam"), native=native)
    rd = lambda n: open(os.path.join(ref_dir, n)).read()  # noqa: E731
    assert O.canonical_bed(out["circ"]) == O.canonical_bed(rd("circ_splice_sites.bed"))
    assert O.canonical_bed(out["lin"]) == O.canonical_bed(rd("lin_splice_sites.bed"))
    assert out["reads"] == rd("spliced_reads.fastq")
    assert O.canonical_multi(out["multi"]) == O.canonical_multi(rd("multi_events.tsv"))
    assert out["counters"] == rd("counters.txt")
