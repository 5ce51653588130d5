#!/usr/bin/env python
"""analyze_results.py — rebuild the comparison CSV/JSON (and PNG, when matplotlib is present) reports from the
``*_parallel_logs/combined_results_{N}.csv`` files of earlier runs, without re-training (the reference README,
README.md:21, refers to this script but the repository never shipped it)."""
import sys

from horizonml_b200.bench_suite import analyze_main

if __name__ == "__main__":
    sys.exit(analyze_main())
