"""ZeRO-1: optimizer-state sharding over the data-parallel group.

The reference replicates Adam's state on every worker (SURVEY §2.3 "ZeRO / FSDP / optimizer sharding: NO").  Two
implementations on top of the flat parameter store:

* ``FusedShardedAdam`` — per gradient bucket ONE peer-memory kernel (``csrc/comm.cu`` ``zero1_kernel``): the two-shot
  all-reduce with the optimizer in the middle.  Every rank packs its bucket (1/W scale, bf16), the owner of each slice sums
  the W copies, applies Adam to its fp32 master slice and its moment shards (2·P/W floats instead of 2·P) and pushes the new
  bf16 parameters into every rank's shadow over NVLink.  Launched by the gradient reducer as buckets complete, i.e.
  overlapped with backward inside the step's CUDA graph; no library collective.  Default for ``--zero1`` where the peer
  kernels run (CUDA, native backend, bf16).  On CPU / gloo the same class runs the same sharding on ``torch.distributed``.
* ``ShardedFlatAdam`` — the baseline formulation: the whole flat gradient buffer is reduce-scattered (NCCL; on gloo an
  all-reduce + slice), the fused Adam kernel runs on the rank's contiguous 1/W slice, the fp32 master slice is all-gathered
  and the bf16 shadow refreshed (``--zero1_impl nccl``).
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


def _peer_device(device: torch.device) -> bool:
    """the peer-memory kernels run there (seam for the CPU dry run of the GPU tests, tests/conftest.py)"""
    return device.type == "cuda"


class FusedShardedAdam:
    """ZeRO-1 on the peer-memory kernel (``csrc/comm.cu`` ``zero1_kernel``): per gradient bucket ONE kernel reduces the
    bucket over the ranks, applies Adam to the slice this rank owns (moment shards, fp32 master authoritative on the
    owner) and pushes the new bf16 parameters into every rank's shadow — launched by the gradient reducer the moment the
    bucket's last gradient is enqueued, i.e. overlapped with the rest of backward, inside the step's CUDA graph.  No
    NCCL reduce-scatter / all-gather, no separate optimizer or shadow-refresh pass.

    Sharding is per bucket in *wire* order (the compacted live elements of the bucket, 8-element vectors): rank r owns
    vectors ``[r·q, (r+1)·q)`` with ``q = ceil(nv / W)``; ``m`` / ``v`` (/ the divergence metric's previous gradient)
    hold ``q·8`` elements per bucket.  ``gather_state()`` reassembles FlatAdam-compatible full tensors (and refreshes
    the fp32 master everywhere) for checkpoints; ``load_state_dict()`` takes the same format.

    On CPU / gloo (no peer memory, fp32 compute) the same sharding runs on ``torch.distributed`` collectives — the
    plumbing path the equivalence tests use; there the fp32 master itself is all-gathered."""

    def __init__(self, flat: FlatParams, ar=None, group=None, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 with_prev: bool = False, early_blocks: int = 24):
        self.flat, self.ar, self.group, self.lr, self.betas, self.eps = flat, ar, group, lr, betas, eps
        ready = dist.is_available() and dist.is_initialized()
        self.world = ar.world if ar is not None else (dist.get_world_size(group) if ready else 1)
        self.rank = ar.rank if ar is not None else (dist.get_rank(group) if ready else 0)
        self.native = ar is not None and hasattr(ar, "handle") and _peer_device(flat.device) and flat.shadow is not None
        dev, W = flat.device, self.world
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.early_blocks = early_blocks
        self.idx: list = []          # per bucket: flat offsets of the bucket's wire elements (int64, on the device)
        self.own: list = []          # per bucket: (first wire element, count) owned by this rank
        self.m, self.v, self.prev = [], [], []
        for b in flat.buckets:
            live = flat.bucket_live[b.index]
            if live is None:
                idx = torch.arange(b.start, b.end, dtype=torch.int64, device=dev)
            else:
                idx = (b.start + live.to(torch.int64)[:, None] * ALIGN + torch.arange(ALIGN, device=dev)[None, :]).reshape(-1)
            n = idx.numel()
            q = (n // 8 + W - 1) // W
            lo = min(self.rank * q, n // 8) * 8
            cnt = max(0, min(lo + q * 8, n) - lo)
            self.idx.append(idx)
            self.own.append((lo, cnt))
            self.m.append(torch.zeros(q * 8, dtype=torch.float32, device=dev))
            self.v.append(torch.zeros(q * 8, dtype=torch.float32, device=dev))
            self.prev.append(torch.zeros(q * 8, dtype=torch.float32, device=dev) if with_prev else None)
        flat.zeroed_by_optimizer = True
        flat.grad.zero_()

    @property
    def state_numel(self) -> int:
        return 2 * sum(t.numel() for t in self.m)

    def state_tensors(self):
        return [self.step_t] + self.m + self.v + [p for p in self.prev if p is not None]

    @torch.no_grad()
    def step_bucket(self, b: int, last: bool, diff_out: Optional[torch.Tensor] = None) -> str:
        """Reduce + sharded Adam + parameter broadcast of bucket ``b``; ``last`` advances the step counter."""
        f, W = self.flat, self.world
        bk = f.buckets[b]
        sl = slice(bk.start, bk.end)
        if self.native:
            self.ar.handle.set_block_cap(0 if last else self.early_blocks)
            self.ar.handle.zero1_step(f.grad[sl], 1.0 / W, f.bucket_live[b], f.master[sl], self.m[b], self.v[b],
                                      f.shadow[sl], self.prev[b], diff_out if self.prev[b] is not None else None,
                                      self.step_t, float(self.lr), float(self.betas[0]), float(self.betas[1]),
                                      float(self.eps), bool(last))
            return "zero1"
        # ---- plumbing path (CPU / gloo, or no peer communicator): same sharding, library collectives
        from .. import ops
        idx, (lo, cnt) = self.idx[b], self.own[b]
        g = f.grad[idx]
        if W > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
            g.mul_(1.0 / W)
        own = idx[lo:lo + cnt]
        if cnt:
            p, gm = f.master[own].contiguous(), g[lo:lo + cnt].contiguous()
            m, v = self.m[b][:cnt], self.v[b][:cnt]
            st = self.step_t.clone()                      # every bucket of a step sees the same step count
            if self.prev[b] is not None and diff_out is not None:
                d = gm - self.prev[b][:cnt]
                part = (d * d).sum()
                self.prev[b][:cnt].copy_(gm)
            ops.adam_step(p, gm, m, v, None, st, self.lr, self.betas[0], self.betas[1], self.eps, 1.0, None, False)
            f.master[own] = p
        if self.prev[b] is not None and diff_out is not None:
            part = part if cnt else torch.zeros((), device=f.device)
            if W > 1:
                dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
            diff_out.add_(part)
        if W > 1:
            self._allgather_into(f.master, b, f.master[own] if cnt else f.master[:0])
        if f.shadow is not None:
            f.shadow[idx] = f.master[idx].to(f.shadow.dtype)
        f.grad[idx] = 0.0
        if last:
            self.step_t += 1.0
        return "zero1-dist"

    # ---- (de)sharding helpers: only used off the hot path (checkpoints, plumbing path)
    def _allgather_into(self, full: torch.Tensor, b: int, mine: torch.Tensor) -> None:
        """full[idx of rank r's slice of bucket b] = rank r's ``mine`` for every r (shards padded to q·8)."""
        W, idx = self.world, self.idx[b]
        shard = self.m[b].numel()
        buf = torch.zeros(shard, dtype=full.dtype, device=full.device)
        buf[:mine.numel()] = mine
        if W == 1:
            full[idx[:mine.numel()]] = mine
            return
        parts = [torch.empty_like(buf) for _ in range(W)]
        dist.all_gather(parts, buf, group=self.group)
        n = idx.numel()
        for r in range(W):
            lo = min(r * shard, n)
            cnt = max(0, min(lo + shard, n) - lo)
            if cnt:
                full[idx[lo:lo + cnt]] = parts[r][:cnt]

    @torch.no_grad()
    def gather_state(self) -> dict:
        """Collective: FlatAdam-compatible full moments on every rank; also makes the fp32 master whole again."""
        f = self.flat
        m = torch.zeros(f.total, dtype=torch.float32, device=f.device)
        v = torch.zeros_like(m)
        for b in range(len(f.buckets)):
            lo, cnt = self.own[b]
            self._allgather_into(m, b, self.m[b][:cnt])
            self._allgather_into(v, b, self.v[b][:cnt])
            if self.native:          # non-owners' master copies are stale on the kernel path
                self._allgather_into(f.master, b, f.master[self.idx[b][lo:lo + cnt]])
        return {"m": m.cpu(), "v": v.cpu(), "step": self.step_t.cpu(), "lr": self.lr, "betas": self.betas,
                "eps": self.eps}

    def state_dict(self) -> dict:
        return self.gather_state()

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        dev = self.flat.device
        fm, fv = sd["m"].to(dev), sd["v"].to(dev)
        for b in range(len(self.flat.buckets)):
            lo, cnt = self.own[b]
            own = self.idx[b][lo:lo + cnt]
            self.m[b].zero_(); self.v[b].zero_()
            self.m[b][:cnt] = fm[own]
            self.v[b][:cnt] = fv[own]
        self.step_t.copy_(sd["step"])
        self.lr = sd.get("lr", self.lr)
