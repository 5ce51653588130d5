"""ZeRO-1: optimizer-state sharding over the data-parallel group.

The reference replicates Adam's state on every worker (SURVEY §2.3 "ZeRO / FSDP / optimizer sharding: NO"); here the
flat parameter store makes the sharded variant a few lines: the flat gradient buffer is reduce-scattered (each rank
receives the average of ITS contiguous 1/W slice), the fused Adam kernel runs on that slice only (moments exist only
for it: 2·P/W floats instead of 2·P), and the updated fp32 master slice is all-gathered; the bf16 shadow is refreshed
from it.  Collectives go through ``torch.distributed`` (NCCL on GPUs; on gloo, which has no reduce-scatter, an
all-reduce + slice) — this is the baseline formulation of the path; folding the Adam update between the two halves
of the two-shot peer all-reduce kernel (csrc/comm.cu) is the fused variant the design points to.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from ..models.flat import ALIGN, FlatParams


class ShardedFlatAdam:
    """Drop-in for ``FlatAdam`` (same ``step`` contract) whose ``step`` also performs the gradient reduction."""

    def __init__(self, flat: FlatParams, group=None, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.flat, self.group, self.lr, self.betas, self.eps = flat, group, lr, betas, eps
        ready = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ready else 1
        self.rank = dist.get_rank(group) if ready else 0
        if flat.total % (ALIGN * self.world):
            raise ValueError("FlatParams must be built with pad_multiple = 64 * world_size for optimizer sharding")
        self.shard = flat.total // self.world
        self.lo, self.hi = self.rank * self.shard, (self.rank + 1) * self.shard
        dev = flat.device
        self.m = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.gshard = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self._gather_src = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.live = None
        if flat.live_blocks is not None:
            idx = flat.live_blocks.cpu()
            lo, hi = self.lo // ALIGN, self.hi // ALIGN
            self.live = (idx[(idx >= lo) & (idx < hi)] - lo).to(torch.int32).to(dev)
        self._native_rs = dev.type == "cuda"          # gloo has no reduce-scatter
        flat.zeroed_by_optimizer = True
        flat.grad.zero_()

    @property
    def state_numel(self) -> int:
        return 2 * self.shard

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, prev_grad: Optional[torch.Tensor] = None):
        """Reduce-scatter(avg) → Adam on the local shard → all-gather parameters.  ``prev_grad`` (shard-sized) keeps
        the reference's gradient-divergence metric; returns its Σ(g−prev)² over ALL shards (or None)."""
        from .. import ops
        f, W = self.flat, self.world
        if W > 1:
            if self._native_rs:
                dist.reduce_scatter_tensor(self.gshard, f.grad, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(f.grad, op=dist.ReduceOp.SUM, group=self.group)
                self.gshard.copy_(f.grad[self.lo:self.hi])
            g = self.gshard.mul_(1.0 / W)            # average (the replicated path folds 1/W into its all-reduce)
        else:
            g = f.grad
        sl = slice(self.lo, self.hi)
        diff = ops.adam_step(f.master[sl], g, self.m, self.v, f.shadow[sl] if f.shadow is not None else None,
                             self.step_t, self.lr, self.betas[0], self.betas[1], self.eps, grad_scale, prev_grad,
                             False, live_blocks=self.live)
        if W > 1:
            self._gather_src.copy_(f.master[sl])
            dist.all_gather_into_tensor(f.master, self._gather_src, group=self.group)
            f.sync_shadow()
            if diff is not None:                     # Σ(g−prev)² of this shard → of the whole gradient
                dist.all_reduce(diff, op=dist.ReduceOp.SUM, group=self.group)
        f.grad.zero_()
        return diff

    # ---- checkpointing: every rank calls gather_state(); it returns the full FlatAdam-compatible state everywhere
    def gather_state(self) -> dict:
        def full(t):
            if self.world == 1:
                return t.detach().cpu().clone()
            out = torch.empty(self.shard * self.world, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
            return out.cpu()
        return {"m": full(self.m), "v": full(self.v), "step": self.step_t.cpu(), "lr": self.lr,
                "betas": self.betas, "eps": self.eps}

    def state_dict(self) -> dict:
        return self.gather_state()

    def load_state_dict(self, sd: dict) -> None:
        self.m.copy_(sd["m"][self.lo:self.hi])
        self.v.copy_(sd["v"][self.lo:self.hi])
        self.step_t.copy_(sd["step"])
        self.lr = sd.get("lr", self.lr)
