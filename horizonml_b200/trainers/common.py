"""Runtime pieces shared by the three trainers: device/dtype/backend resolution, the CUDA-graph
captured step, device-side metric accumulators, heartbeat + fault-injection hooks."""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from .. import ops
from ..config import TrainConfig


@dataclass
class Runtime:
    rank: int
    world: int
    device: torch.device
    dtype: torch.dtype
    backend: str           # op backend actually in use
    comm_backend: str


def setup_runtime(rank: int, world: int, cfg: TrainConfig, device: str) -> Runtime:
    from ..launch import init_distributed
    comm_backend = init_distributed(rank, world, device, cfg.comm)
    dev = torch.device("cuda", torch.cuda.current_device()) if device == "cuda" else torch.device("cpu")
    dtype = {"auto": torch.bfloat16 if dev.type == "cuda" else torch.float32,
             "bf16": torch.bfloat16, "fp32": torch.float32}[cfg.dtype]
    backend = cfg.backend
    if backend == "auto":
        backend = "native" if (dev.type == "cuda" and ops.native_available()) else "torch"
    if backend == "native" and dev.type != "cuda":
        raise RuntimeError("--backend native requires a CUDA device")
    ops.set_backend(backend)
    if dev.type == "cpu":
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    else:
        torch.set_num_threads(1)     # the host only launches kernels: no spinning OpenMP teams next to it
    torch.manual_seed(cfg.seed)
    return Runtime(rank, world, dev, dtype, backend, comm_backend)


class _NullRange:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL_RANGE = _NullRange()
_NVTX = os.environ.get("HZ_NVTX", "0") == "1"


class _NvtxRange:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *a):
        torch.cuda.nvtx.range_pop()
        return False


def nvtx_range(name: str):
    """``with nvtx_range("forward"):`` — an NVTX range around a region of the step when ``HZ_NVTX=1`` (or ``--profile``
    sets it) on a CUDA box, for nsys / ncu ``--nvtx`` filtering; otherwise a shared no-op object, so the
    hot path pays one attribute lookup."""
    if _NVTX and torch.cuda.is_available():
        return _NvtxRange(name)
    return _NULL_RANGE


def enable_nvtx(flag: bool = True) -> None:
    global _NVTX
    _NVTX = bool(flag)


class DeviceStats:
    """On-device accumulators: no ``.item()`` in the step loop (the reference syncs three times per
    step: data_parallel_train.py:126,130,140)."""

    N = 6   # loss_sum, correct, seen, grad_div_sum, grad_div_n, steps

    def __init__(self, device):
        self.buf = torch.zeros(self.N, dtype=torch.float32, device=device)
        self.has_prev = torch.zeros(1, dtype=torch.float32, device=device)

    def add_step(self, loss, correct, batch: int, diff_sq=None) -> None:
        """One tiny kernel: loss/accuracy/sample counters (+ grad-divergence term when given)."""
        ops.stats_update(self.buf, self.has_prev, loss, correct, float(batch), diff_sq)

    def read_and_reset(self, keep_div: bool = True) -> Dict[str, float]:
        v = self.buf.tolist()
        out = {"loss_sum": v[0], "correct": v[1], "seen": v[2], "grad_div_sum": v[3],
               "grad_div_n": v[4], "steps": v[5]}
        self.buf[:3].zero_()
        self.buf[5].zero_()
        if not keep_div:
            self.buf[3:5].zero_()
        return out


class GraphedStep:
    """Runs ``fn(images, labels)`` eagerly for ``warmup`` calls, then captures it into a CUDA graph
    (static input buffers) and replays.  Falls back to eager for off-size batches / CPU."""

    def __init__(self, fn: Callable, device, enabled: bool, warmup: int = 3):
        self.fn, self.device = fn, torch.device(device)
        self.enabled = enabled and self.device.type == "cuda"
        self.warmup = warmup
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_x = self.static_y = None
        self.shape = None
        self.capture_error: Optional[str] = None

    def __call__(self, x, y):
        self.calls += 1
        if not self.enabled:
            return self.fn(x, y)
        if self.graph is not None and tuple(x.shape) == self.shape:
            if x.data_ptr() != self.static_x.data_ptr():      # callers may fill input_buffers() directly
                self.static_x.copy_(x, non_blocking=True)
            if y.data_ptr() != self.static_y.data_ptr():
                self.static_y.copy_(y, non_blocking=True)
            self.graph.replay()
            return None
        if self.graph is None and self.calls > self.warmup and self.capture_error is None:
            try:
                self._capture(x, y)
                self.graph.replay()
                return None
            except Exception as e:  # noqa: BLE001
                self.capture_error = repr(e)
                self.graph = None
                self.enabled = False
                if os.environ.get("HZ_STRICT_GRAPH", "0") == "1":
                    raise
                print(f"[graph] capture failed, staying eager: {e!r}", flush=True)
                torch.cuda.synchronize()
        return self.fn(x, y)

    def input_buffers(self):
        """(images, labels) device buffers the captured graph reads, or None before capture: a loader can land its
        host→device copy straight in them (stream-ordered behind the previous replay) and skip the staging copy."""
        if self.graph is None:
            return None
        return self.static_x, self.static_y

    def _capture(self, x, y):
        self.static_x = x.clone()
        self.static_y = y.clone()
        self.shape = tuple(x.shape)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.fn(self.static_x, self.static_y)
        self.graph = g


class Heartbeat:
    """Per-rank heartbeat file (failure detection aid; the reference has only timeouts)."""

    def __init__(self, directory: Optional[str], rank: int):
        self.path = os.path.join(directory, f"heartbeat_rank{rank}.txt") if directory else None
        if self.path:
            os.makedirs(directory, exist_ok=True)

    def beat(self, epoch: int, step: int) -> None:
        if self.path:
            with open(self.path, "w") as fh:
                fh.write(f"{time.time():.3f} epoch={epoch} step={step}\n")


class FaultInjector:
    """``--inject_fault rank:step`` — that rank dies at that global step (launcher tear-down test)."""

    def __init__(self, spec: Optional[str], rank: int):
        self.step = None
        if spec:
            r, s = spec.split(":")
            if int(r) == rank:
                self.step = int(s)

    def maybe_fail(self, global_step: int) -> None:
        if self.step is not None and global_step >= self.step:
            raise RuntimeError(f"injected fault at global step {global_step}")


def gpu_mem_mb(device) -> float:
    if torch.device(device).type != "cuda":
        return 0.0
    return torch.cuda.max_memory_allocated(device) / (1 << 20)


def allreduce_max_scalar(x: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device if torch.device(device).type == "cuda" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def profile_steps(step_fn, x, y, path: str, steps: int = 5) -> None:
    """``--profile``: CUPTI kernel timeline of a few steps (chrome trace + per-kernel summary JSON)."""
    import json
    from torch.profiler import ProfilerActivity, profile
    acts = [ProfilerActivity.CPU] + ([ProfilerActivity.CUDA] if torch.cuda.is_available() else [])
    with profile(activities=acts) as prof:
        for _ in range(steps):
            step_fn(x, y)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    prof.export_chrome_trace(path)
    rows = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(e.name[:80], [0, 0.0])
            r[0] += 1
            r[1] += e.time_range.end - e.time_range.start
    with open(path.replace(".json", "_kernels.json"), "w") as fh:
        json.dump({k: {"launches": v[0], "us": v[1]} for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])}, fh, indent=1)
