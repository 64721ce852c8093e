#!/usr/bin/env python3
"""drop-in entry point with the reference's name: see find_circ2_b200/cli.py"""
import sys

from find_circ2_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
