#!/usr/bin/env python
"""main.py — benchmark sweep over the three strategies + comparison reports
(reference main.py: --sample_sizes 1000 10000 50000 --world_size 5 --epochs 5 --output_dir benchmark_results)."""
import sys

from horizonml_b200.bench_suite import generate_comparison_graphs, main, run_benchmarks  # noqa: F401

if __name__ == "__main__":
    sys.exit(main())
