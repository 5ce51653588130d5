"""Flat parameter / gradient storage.

All parameters of a (sub)model live in ONE fp32 master buffer, their gradients in ONE fp32
gradient buffer laid out **in reverse execution order** (fc, layer4 … conv1 — the order backward
produces them, SURVEY §2.5 "Backward shapes"), and — for bf16 compute — a bf16 shadow buffer the
kernels read and the fused Adam kernel refreshes.  Buckets are contiguous slices of the gradient
buffer, so a bucket all-reduce needs no flatten/unflatten copy (what DDP's reducer does with
25 MiB buckets, data_parallel_train.py:202) and wgrad kernels write straight into it.

Every parameter's offset is padded to 64 elements so bf16 shadows are 128-byte aligned (TMA global
address requirement is 16 B; 128 B keeps vector loads and swizzle atoms aligned).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

ALIGN = 64
# bucket-wise Adam runs beside backward kernels: 2 CTAs x 256 threads per SM leave room for their CTAs
_BUCKET_CTAS = int(os.environ.get("HZ_ADAM_BUCKET_CTAS", "296"))


@dataclass
class Bucket:
    index: int
    start: int           # element offset in the flat buffers
    end: int
    names: List[str]


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class FlatParams:
    def __init__(self, named_params: Sequence[Tuple[str, nn.Parameter]], device, compute_dtype,
                 bucket_cap_mb: float = 25.0, reverse: bool = True, first_bucket_mb: float = 1.0,
                 live_masks: Optional[Dict[str, torch.Tensor]] = None, bucket_by_live: bool = False,
                 pad_multiple: int = ALIGN, bucket_starts: Optional[Sequence[str]] = None):
        named = list(named_params)
        order = list(reversed(named)) if reverse else named
        self.names = [n for n, _ in order]
        self.params = [p for _, p in order]
        offs, cur = [], 0
        for p in self.params:
            offs.append(cur)
            cur = _round_up(cur + p.numel(), ALIGN)
        self.offsets = offs
        # pad_multiple = 64 * world lets an optimizer shard the flat buffers evenly (parallel/zero.py)
        self.total = max(_round_up(cur, max(pad_multiple, ALIGN)), ALIGN)
        self.device = torch.device(device)
        self.compute_dtype = compute_dtype
        self.master = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.shadow = (torch.zeros(self.total, dtype=compute_dtype, device=self.device)
                       if compute_dtype != torch.float32 else None)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            src = p.detach()
            cl = src.dim() == 4
            # physical order of a channels_last 4-D tensor is [d0, d2, d3, d1]
            phys = src.permute(0, 2, 3, 1).contiguous().view(-1) if cl else src.contiguous().view(-1)
            self.master[o:o + n].copy_(phys)

            def view(buf):
                v = buf[o:o + n]
                if cl:
                    d0, d1, d2, d3 = src.shape
                    return v.view(d0, d2, d3, d1).permute(0, 3, 1, 2)
                return v.view(src.shape)

            p.data = view(self.master)
            p.main_grad = view(self.grad)
            p._acc = False
            p._flat_range = (o, o + n)
            if self.shadow is not None:
                p.shadow = view(self.shadow)
                p.shadow._hz_stable = True      # only the optimizer pass writes it (see native_backend._stable)
        self.sync_shadow()
        self._attach_autograd_bridge()
        self.live_blocks = None               # int32 indices of the 64-element blocks that can be non-zero
        blk = self._live_block_mask(live_masks) if live_masks else None
        # bucket_by_live: the caps count elements that actually travel (dead taps excluded), so a "25 MiB" bucket
        # of mostly-dead layer4 weights does not delay the collective of the live ones behind it
        # bucket_starts: explicit boundaries — a parameter whose name starts with one of these prefixes opens a new
        # bucket the first time the prefix is met (gradient-ready order), e.g. ("layer3.", "layer2.", "layer1.") gives
        # [fc+layer4] [layer3] [layer2] [layer1+stem]: every bucket's collective starts the moment its layer group's
        # backward is done and the bucket that can only be reduced after the very last gradient is the smallest one
        self.buckets = (self._make_buckets_at(bucket_starts) if bucket_starts else
                        self._make_buckets(bucket_cap_mb, first_bucket_mb, blk if bucket_by_live else None))
        self.bucket_live: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        if blk is not None:
            self._build_live(blk)
        self._bucket_of: Dict[int, int] = {}
        for b in self.buckets:
            for nme in b.names:
                self._bucket_of[id(self.params[self.names.index(nme)])] = b.index

    # ------------------------------------------------------------------------------------
    def _make_buckets(self, cap_mb: float, first_mb: float, live_blk: Optional[torch.Tensor] = None) -> List[Bucket]:
        """Greedy contiguous bucketing in gradient-ready order (DDP-style: small first bucket so
        the first all-reduce starts early, then ``cap_mb`` buckets).  With ``live_blk`` (bool per 64-element
        block) sizes are measured in live elements."""
        buckets: List[Bucket] = []
        cap = int(first_mb * (1 << 20) / 4)
        start, names = 0, []
        csum = None
        if live_blk is not None:
            csum = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(live_blk.long(), 0)]) * ALIGN
        for i, (nme, p, o) in enumerate(zip(self.names, self.params, self.offsets)):
            names.append(nme)
            end = self.offsets[i + 1] if i + 1 < len(self.offsets) else self.total
            size = end - start if csum is None else int(csum[end // ALIGN] - csum[start // ALIGN])
            if size >= cap or i + 1 == len(self.params):
                buckets.append(Bucket(len(buckets), start, end, names))
                start, names = end, []
                cap = int(cap_mb * (1 << 20) / 4)
        return buckets

    def _make_buckets_at(self, starts: Sequence[str]) -> List[Bucket]:
        buckets: List[Bucket] = []
        seen = set()
        start, names = 0, []
        for i, nme in enumerate(self.names):
            hit = next((s for s in starts if nme.startswith(s) and s not in seen), None)
            if hit is not None:
                seen.add(hit)
                if names:
                    buckets.append(Bucket(len(buckets), start, self.offsets[i], names))
                    start, names = self.offsets[i], []
            names.append(nme)
        buckets.append(Bucket(len(buckets), start, self.total, names))
        return buckets

    def _live_block_mask(self, live_masks) -> Optional[torch.Tensor]:
        """bool per 64-element block: can any element of it ever receive a gradient?  None when all can."""
        nblk = self.total // ALIGN
        live = torch.ones(self.total, dtype=torch.bool)
        for nme, p, o in zip(self.names, self.params, self.offsets):
            m = live_masks.get(nme)
            if m is None:
                continue
            phys = m.permute(0, 2, 3, 1).reshape(-1) if m.dim() == 4 else m.reshape(-1)
            live[o:o + p.numel()] = phys
        blk = live.view(nblk, ALIGN).any(dim=1)
        return None if bool(blk.all()) else blk

    def _build_live(self, blk: torch.Tensor) -> None:
        """Dead-parameter elision: block list for the optimizer (whole buffer) and per bucket (all-reduce)."""
        nblk = self.total // ALIGN
        idx = torch.nonzero(blk).flatten().to(torch.int32)
        self.live_blocks = idx.to(self.device)
        self.live_fraction = float(idx.numel()) / nblk
        for b in self.buckets:
            lo, hi = b.start // ALIGN, b.end // ALIGN
            sel = idx[(idx >= lo) & (idx < hi)] - lo
            if sel.numel() < hi - lo:
                self.bucket_live[b.index] = sel.to(torch.int32).to(self.device)

    def _attach_autograd_bridge(self) -> None:
        """Parameters used by plain autograd ops (library models, e.g. MobileNetV2) receive ``.grad``;
        bridge it into ``main_grad`` so the reducer/optimizer see one storage."""
        def bridge(p):
            g = p.grad
            if g is None:
                return
            if getattr(p, "_acc", False):
                p.main_grad.add_(g.to(torch.float32))
            else:
                p.main_grad.copy_(g.to(torch.float32))
            p.grad = None
            p._acc = True
            hook = getattr(p, "_ready_hook", None)
            if hook is not None:
                hook(p)

        for p in self.params:
            if p.requires_grad:
                p.register_post_accumulate_grad_hook(bridge)

    def bucket_index(self, p) -> int:
        return self._bucket_of[id(p)]

    def begin_step(self) -> None:
        """Start of an optimizer step: the first gradient write of each parameter overwrites, later ones
        (micro-batches) accumulate.  When the optimizer clears the buffer in its own pass
        (``zeroed_by_optimizer``) producers that add atomically need no memset either."""
        zeroed = bool(getattr(self, "zeroed_by_optimizer", False))
        for p in self.params:
            p._acc = False          # first write of the step may overwrite ...
            p._zeroed = zeroed      # ... and split-K / atomic producers may rely on an all-zero buffer

    def zero_grad(self) -> None:
        self.grad.zero_()
        self.begin_step()

    @torch.no_grad()
    def sync_shadow(self) -> None:
        if self.shadow is not None:
            self.shadow.copy_(self.master)

    def numel(self) -> int:
        return sum(p.numel() for p in self.params)

    def grad_bytes(self, dtype_bytes: int = 4) -> int:
        return self.numel() * dtype_bytes


class FlatAdam:
    """Adam over a FlatParams store — one fused kernel per step, fp32 master +
    moments, bf16 shadow refresh fused in.  Semantics = ``torch.optim.Adam(lr)`` as used by the
    reference (data_parallel_train.py:205)."""

    def __init__(self, flat: FlatParams, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.flat, self.lr, self.betas, self.eps = flat, lr, betas, eps
        self.m = torch.zeros_like(flat.master)
        self.v = torch.zeros_like(flat.master)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=flat.device)
        # the Adam pass also clears the gradient buffer → producers only ever accumulate, no memsets
        flat.zeroed_by_optimizer = True
        flat.grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0, prev_grad=None):
        """One fused pass: Adam + bf16 shadow refresh + (optional) gradient-divergence Σ(g−prev)² +
        gradient clear.  Returns the divergence term (0-d tensor) or None."""
        from .. import ops
        f = self.flat
        return ops.adam_step(f.master, f.grad, self.m, self.v, f.shadow, self.step_t, self.lr,
                             self.betas[0], self.betas[1], self.eps, grad_scale, prev_grad, True,
                             live_blocks=f.live_blocks)

    @torch.no_grad()
    def step_bucket(self, b: int, first: bool, diff_out=None, prev_grad=None, grad_scale: float = 1.0) -> None:
        """The same fused pass restricted to bucket ``b`` — launched by the gradient reducer the moment that
        bucket's gradients are final, so the optimizer overlaps the rest of backward.  ``first`` marks the first
        bucket processed in this step (advances the step counter, clears the divergence accumulator)."""
        from .. import ops
        f = self.flat
        bk = f.buckets[b]
        sl = slice(bk.start, bk.end)
        ops.adam_step(f.master[sl], f.grad[sl], self.m[sl], self.v[sl],
                      f.shadow[sl] if f.shadow is not None else None, self.step_t, self.lr,
                      self.betas[0], self.betas[1], self.eps, grad_scale,
                      prev_grad[sl] if prev_grad is not None else None, True,
                      live_blocks=f.bucket_live[b], diff_out=diff_out if prev_grad is not None else None, bump=first,
                      max_ctas=_BUCKET_CTAS)

    def state_dict(self) -> dict:
        return {"m": self.m.cpu(), "v": self.v.cpu(), "step": self.step_t.cpu(), "lr": self.lr,
                "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd: dict) -> None:
        self.m.copy_(sd["m"]); self.v.copy_(sd["v"]); self.step_t.copy_(sd["step"])
        self.lr = sd.get("lr", self.lr)
