#!/usr/bin/env python
"""tensor_parallel_train.py — same CLI as the reference script of the same name
(--world_size 5 --epochs 5 --sample_size 1000); implementation in horizonml_b200."""
import sys

from horizonml_b200.cli import tensor_parallel_main
from horizonml_b200.trainers import run_data_parallel, run_model_parallel, run_tensor_parallel  # noqa: F401

if __name__ == "__main__":
    sys.exit(tensor_parallel_main())
