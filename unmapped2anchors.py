#!/usr/bin/env python3
"""drop-in for the reference's unmapped2anchors.py (see find_circ2_b200/unmapped2anchors.py)"""
import sys

from find_circ2_b200.unmapped2anchors import main

if __name__ == "__main__":
    sys.exit(main())
