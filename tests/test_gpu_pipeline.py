"""
End-to-end drop-in parity on the GPU: the product pipeline (host logic + CUDA scan + CUDA aggregation) must
reproduce every output file the REFERENCE produced (tests/golden/*/ref_*) byte for byte after the canonical sort,
for every stored option set.
"""
import os

import pytest

from conftest import golden_cases, golden_ids
import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case_dir,ref_dir,argv", golden_cases(), ids=golden_ids())
@pytest.mark.parametrize("batch,native", [(1 << 18, True), (97, False), (1500, True)])
def test_pipeline_matches_reference(case_dir, ref_dir, argv, batch, native):
    from find_circ2_b200 import cli

    opt = cli.parse_args(["-G", os.path.join(case_dir, "genome.fa")] + argv + [os.path.join(case_dir, "input.sam")])[0]
    opt.batch_pairs = batch
    if native and (opt.allhits or opt.noop or opt.test):
        native = False  # --all-hits, --noop and --test always use the python ingest
    out = cli.run_to_strings(opt, os.path.join(case_dir, "input.sam"), native=native)
    H.compare_outputs(out["circ"], out["lin"], out["reads"], out["multi"], out["counters"], ref_dir, argv, out["test"])
