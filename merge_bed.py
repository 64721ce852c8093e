#!/usr/bin/env python3
"""drop-in for the reference's merge_bed.py (see find_circ2_b200/merge_bed.py)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from find_circ2_b200.merge_bed import main  # noqa: E402

sys.exit(main())
