import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    """[(case_dir, ref_dir, argv)] for every reference run stored under tests/golden/"""
    out = []
    for case in sorted(os.listdir(GOLDEN)):
        cdir = os.path.join(GOLDEN, case)
        if not os.path.isdir(cdir) or not os.path.exists(os.path.join(cdir, "input.sam")):
            continue
        for ref in sorted(os.listdir(cdir)):
            rdir = os.path.join(cdir, ref)
            if not ref.startswith("ref_") or not os.path.isdir(rdir):
                continue
            lines = open(os.path.join(rdir, "cmdline.txt")).read().split("\n")
            if lines[1].strip() != "exit=0":
                continue
            out.append((cdir, rdir, [a.replace("@CASE@", cdir) for a in lines[0].split()]))
    return out


def golden_ids():
    return ["%s-%s" % (os.path.basename(c), os.path.basename(r)[4:]) for c, r, _ in golden_cases()]
