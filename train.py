#!/usr/bin/env python
"""train.py — legacy one-container-per-rank data-parallel entrypoint (reference train.py:15-126).

Rendezvous comes from the environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT), the model from
MODEL_TYPE ∈ {resnet, mobilenet}; writes ``training_logs_worker_{rank}.csv`` with the reference's
columns ``Worker,Epoch,Loss,Accuracy,Time`` (train.py:115-116).  Extra env knobs: EPOCHS (5),
BATCH_SIZE (reference literal 2), SAMPLE_SIZE (all), DEVICE (auto).
"""
import os
import sys

import pandas as pd

from horizonml_b200.config import TrainConfig
from horizonml_b200.launch import resolve_device
from horizonml_b200.trainers.dp import train_data_parallel


def main() -> int:
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    model = {"resnet": "resnet18", "mobilenet": "mobilenet"}[os.environ.get("MODEL_TYPE", "resnet")]
    cfg = TrainConfig(strategy="data", world_size=world, epochs=int(os.environ.get("EPOCHS", 5)),
                      sample_size=int(os.environ.get("SAMPLE_SIZE", 50000)),
                      batch_size=int(os.environ.get("BATCH_SIZE", 2)), model=model,
                      device=os.environ.get("DEVICE", "auto"), logs_dir=os.environ.get("LOGS_DIR", "."),
                      synthetic=os.environ.get("REAL_DATA", "0") != "1")
    df = train_data_parallel(rank, world, cfg, resolve_device(cfg))
    legacy = pd.DataFrame({"Worker": rank, "Epoch": df["epoch"], "Loss": df["loss"],
                           "Accuracy": df["accuracy"], "Time": df["epoch_time"]})
    legacy.to_csv(f"training_logs_worker_{rank}.csv", index=False)
    return 0


if __name__ == "__main__":
    sys.exit(main())
