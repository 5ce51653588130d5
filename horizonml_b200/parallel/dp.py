"""Data-parallel gradient reducer.

Reference: stock ``DistributedDataParallel`` on gloo (data_parallel_train.py:202) — ≈3 buckets
(9.0 / 25.3 / 8.4 MiB fp32) all-reduced from autograd hooks during ``loss.backward()``, then a separate
``optimizer.step()`` (:121).

Here: gradients already live in contiguous reverse-order buckets (``FlatParams``); every wgrad /
BN-backward kernel calls ``param._ready_hook`` after enqueueing its write, and when the last
parameter of a bucket is ready the bucket's fused all-reduce kernel is enqueued on a dedicated
**comm stream** behind an event, overlapping with the remaining backward kernels on the compute
stream.  An optional ``post_bucket(b, first)`` callback runs on the same stream right behind the
collective — the trainers pass the bucket-wise fused Adam, so the optimizer pass (HBM-bound) hides
under the rest of backward (latency-bound) too, also on a single GPU.  ``finish()`` joins the
streams.  The whole thing is CUDA-graph capturable (fork/join through events).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from ..models.flat import FlatParams
from .comm import GradAllReduce


class GradReducer:
    def __init__(self, flat: FlatParams, allreduce: Optional[GradAllReduce], overlap: bool = True,
                 post_bucket: Optional[Callable[[int, bool], None]] = None,
                 fused_bucket: Optional[Callable[[int, bool], None]] = None):
        """``fused_bucket(b, last)`` replaces all-reduce + ``post_bucket`` by ONE kernel per bucket that reduces the
        gradients over the peers and applies the optimizer to the bucket's parameters (csrc/comm.cu ``AdamFuse``)."""
        self.flat, self.ar = flat, allreduce
        self.world = allreduce.world if allreduce is not None else 1
        self.post_bucket = post_bucket
        self.fused_bucket = fused_bucket
        self.algos: List[Optional[str]] = [None] * len(flat.buckets)
        self.cuda = flat.device.type == "cuda"
        # NCCL collectives are captured on the main stream (the well-trodden CUDA-graph path); our peer kernels
        # (and, on one GPU, the bucket-wise optimizer alone) overlap with backward on a dedicated comm stream
        self.overlap = (overlap and self.cuda and (self.world > 1 or post_bucket is not None)
                        and (allreduce is None or allreduce.name != "nccl"))
        self.enabled = True
        self.comm_stream = torch.cuda.Stream(device=flat.device) if self.overlap else None
        self._pending: List[int] = []
        self._main_stream = None
        self._counts = [len(b.names) for b in flat.buckets]
        self._launched: List[bool] = []
        self._events: List[Optional[torch.cuda.Event]] = []
        self._n_launched = 0
        self.bytes_last_step = 0
        for p in flat.params:
            p._ready_hook = self._on_ready
        self.begin_step()

    @property
    def active(self) -> bool:
        return self.enabled and (self.world > 1 or self.post_bucket is not None)

    def begin_step(self) -> None:
        self._main_stream = torch.cuda.current_stream(self.flat.device) if self.cuda else None
        self._pending = list(self._counts)
        self._launched = [False] * len(self._counts)
        self._events = [None] * len(self._counts)
        self._n_launched = 0
        self.bytes_last_step = 0
        self._seen = set()

    # called from inside backward, right after the kernel producing p's gradient was enqueued
    def _on_ready(self, p) -> None:
        if not self.active:
            return
        b = self.flat.bucket_index(p)
        if id(p) in self._seen:          # one ready-hook per parameter per step: a repeat (module applied twice,
            return                       # gradient accumulation) must not release the bucket early
        self._seen.add(id(p))
        self._pending[b] -= 1
        # without overlap every bucket is launched from finish() (after the side stream has been joined)
        if self.overlap and self._pending[b] == 0 and not self._launched[b]:
            self._launch(b)

    def _work(self, b: int) -> None:
        bk = self.flat.buckets[b]
        if self.fused_bucket is not None and self.world > 1:
            live = self.flat.bucket_live[b]
            self.bytes_last_step += self.ar.wire_bytes(bk.end - bk.start if live is None else live.numel() * 64)
            # buckets become ready in index order (reverse execution order), so the highest index is the last one
            self.algos[b] = self.fused_bucket(b, self._n_launched == len(self.flat.buckets) - 1)
            self._n_launched += 1
            return
        if self.world > 1:
            live = self.flat.bucket_live[b] if hasattr(self.flat, "bucket_live") else None
            self.bytes_last_step += self.ar.wire_bytes(bk.end - bk.start if live is None else live.numel() * 64)
            if self.overlap and hasattr(self.ar, "plain_blocks"):
                # every bucket but the last is reduced while backward still runs: few CTAs, the SMs stay with backward
                self.ar.allreduce_avg_(self.flat.grad[bk.start:bk.end], live=live,
                                       background=self._n_launched < len(self.flat.buckets) - 1)
            else:
                self.ar.allreduce_avg_(self.flat.grad[bk.start:bk.end], live=live)
        if self.post_bucket is not None:
            self.post_bucket(b, self._n_launched == 0)
        self._n_launched += 1

    def _launch(self, b: int) -> None:
        self._launched[b] = True
        if self.overlap:
            # gradients of one bucket are produced on two streams: BN/dgrad chain (main) and the wgrad
            # side stream — the collective / optimizer must wait for both
            from ..ops import functional as F
            streams = {torch.cuda.current_stream(self.flat.device)}
            if F._side["stream"] is not None:
                streams.add(F._side["stream"])
                streams.add(self._main_stream or torch.cuda.current_stream(self.flat.device))
            for st in streams:
                ready = torch.cuda.Event()
                ready.record(st)
                self.comm_stream.wait_event(ready)
            with torch.cuda.stream(self.comm_stream):
                self._work(b)
                done = torch.cuda.Event()
                done.record(self.comm_stream)
            self._events[b] = done
        else:
            self._work(b)

    def finish(self) -> None:
        """Flush buckets that never filled (e.g. unused params) and join the comm stream."""
        if not self.active:
            return
        for b, launched in enumerate(self._launched):
            if not launched:
                self._launch(b)
        if self.overlap:
            cur = torch.cuda.current_stream(self.flat.device)
            for ev in self._events:
                if ev is not None:
                    cur.wait_event(ev)
