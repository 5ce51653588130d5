"""Checkpoint / resume (absent in the reference — SURVEY §5.4).

Strategy-aware sharding: DP writes one file from rank 0 (replicas are identical); PP writes one
file per stage; TP writes the replicated part once (rank 0) plus one shard file per rank.
Each file holds model state, flat Adam moments, epoch/step counters and RNG state.
"""
from __future__ import annotations

import os
from typing import Optional

import torch


def _path(save_dir: str, tag: str) -> str:
    return os.path.join(save_dir, f"ckpt_{tag}.pt")


def save(save_dir: str, tag: str, model, optimizer, epoch: int, global_step: int, extra: Optional[dict] = None,
         optim_state: Optional[dict] = None):
    """``optim_state``: an already-collected optimizer state (sharded optimizers gather it collectively first)."""
    os.makedirs(save_dir, exist_ok=True)
    sd = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}
    payload = {
        "model": sd,
        "optim": optim_state if optim_state is not None else (optimizer.state_dict() if optimizer is not None else None),
        "flat_names": list(optimizer.flat.names) if optimizer is not None else None,
        "epoch": epoch, "global_step": global_step,
        "rng": torch.random.get_rng_state(),
        "extra": extra or {},
    }
    tmp = _path(save_dir, tag) + ".tmp"
    torch.save(payload, tmp)
    os.replace(tmp, _path(save_dir, tag))
    return _path(save_dir, tag)


def load(save_dir: str, tag: str, model, optimizer) -> Optional[dict]:
    p = _path(save_dir, tag)
    if not os.path.exists(p):
        return None
    payload = torch.load(p, map_location="cpu", weights_only=False)
    own = model.state_dict()
    with torch.no_grad():
        for k, v in payload["model"].items():
            if k in own:
                own[k].copy_(v.to(own[k].device))
    if optimizer is not None and payload.get("optim") is not None:
        optimizer.load_state_dict(payload["optim"])
        optimizer.flat.sync_shadow()
    torch.random.set_rng_state(payload["rng"])
    return payload
