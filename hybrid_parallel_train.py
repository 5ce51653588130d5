#!/usr/bin/env python
"""hybrid_parallel_train.py — the "hybrid" composition the reference README's tagline promises (README.md:5) but
none of its scripts implements: a DP × PP or DP × TP process mesh.

    python hybrid_parallel_train.py --world_size 4 --dp_replicas 2 --inner layer     # 2 replicas of a 2-stage pipeline
    python hybrid_parallel_train.py --world_size 8 --dp_replicas 2 --inner tensor    # 2 replicas of a TP-4 group

Same flags / CSV artefacts as layer_model_parallel_train.py / tensor_parallel_train.py (logs under
``hybrid_<inner>_logs``); every replica trains on its own shard of the data and gradients are averaged over the
data-parallel group of each stage / shard."""
import argparse
import sys

from horizonml_b200.config import add_train_flags, config_from_args
from horizonml_b200.launch import run_strategy

_FN = {"layer": "horizonml_b200.trainers.pp:train_model_parallel",
       "tensor": "horizonml_b200.trainers.tp:train_tensor_parallel"}


def main(argv=None) -> int:
    pre = argparse.ArgumentParser(add_help=False)
    pre.add_argument("--inner", default="layer", choices=["layer", "tensor"])
    known, rest = pre.parse_known_args(argv)
    p = argparse.ArgumentParser(description="Hybrid (data x layer | data x tensor) parallel training", parents=[pre])
    add_train_flags(p, known.inner)
    p.set_defaults(world_size=4, dp_replicas=2)
    args = p.parse_args(argv)
    cfg = config_from_args(args, args.inner)
    if cfg.dp_replicas < 1 or cfg.world_size % cfg.dp_replicas:
        p.error(f"--dp_replicas {cfg.dp_replicas} must divide --world_size {cfg.world_size}")
    if cfg.logs_dir is None:
        cfg = cfg.replace(logs_dir=f"hybrid_{args.inner}_logs")
    df = run_strategy(cfg, _FN[args.inner])
    return 0 if df is not None else 1


if __name__ == "__main__":
    sys.exit(main())
