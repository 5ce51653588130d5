"""CLI entrypoints (reference __main__ blocks: data_parallel_train.py:293-300,
layer_model_parallel_train.py:425-432, tensor_parallel_train.py:387-394, main.py:392-408)."""
from __future__ import annotations

import argparse
import sys

from .config import add_train_flags, config_from_args

_DESC = {"data": "Data Parallel Training", "layer": "Model Parallel Training",
         "tensor": "Tensor Parallel Training"}
_FN = {"data": "horizonml_b200.trainers.dp:train_data_parallel",
       "layer": "horizonml_b200.trainers.pp:train_model_parallel",
       "tensor": "horizonml_b200.trainers.tp:train_tensor_parallel"}


def trainer_main(strategy: str, argv=None) -> int:
    from .launch import run_strategy
    p = argparse.ArgumentParser(description=_DESC[strategy])
    add_train_flags(p, strategy)
    args = p.parse_args(argv)
    cfg = config_from_args(args, strategy)
    df = run_strategy(cfg, _FN[strategy])
    import os
    under_torchrun = "RANK" in os.environ and int(os.environ.get("RANK", "0")) != 0
    return 0 if (df is not None or under_torchrun) else 1


def data_parallel_main(argv=None) -> int:
    return trainer_main("data", argv)


def layer_parallel_main(argv=None) -> int:
    return trainer_main("layer", argv)


def tensor_parallel_main(argv=None) -> int:
    return trainer_main("tensor", argv)
