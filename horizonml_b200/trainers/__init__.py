"""Per-strategy trainers and their ``run_*`` launch functions (the reference's programmatic API,
main.py:8-10: ``run_data_parallel / run_model_parallel / run_tensor_parallel``)."""
from __future__ import annotations

from typing import Optional

from ..config import TrainConfig
from ..launch import run_strategy


def _cfg(strategy, world_size, epochs, sample_size, logs_dir, kw) -> TrainConfig:
    cfg = kw.pop("cfg", None) or TrainConfig()
    return cfg.replace(strategy=strategy, world_size=world_size, epochs=epochs,
                       sample_size=sample_size, logs_dir=logs_dir, **kw)


def run_data_parallel(world_size, epochs, sample_size, logs_dir="data_parallel_logs", **kw):
    return run_strategy(_cfg("data", world_size, epochs, sample_size, logs_dir, kw),
                        "horizonml_b200.trainers.dp:train_data_parallel")


def run_model_parallel(world_size, epochs, sample_size, logs_dir="model_parallel_logs", **kw):
    return run_strategy(_cfg("layer", world_size, epochs, sample_size, logs_dir, kw),
                        "horizonml_b200.trainers.pp:train_model_parallel")


def run_tensor_parallel(world_size, epochs, sample_size, logs_dir="tensor_parallel_logs", **kw):
    return run_strategy(_cfg("tensor", world_size, epochs, sample_size, logs_dir, kw),
                        "horizonml_b200.trainers.tp:train_tensor_parallel")
