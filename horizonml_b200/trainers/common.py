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


def split_compute_comm(dev_s: float, regions: Dict[str, float]):
    """(compute_s, comm_s, source) of an epoch's device time in the reference's convention (its
    "compute_time" stopwatch brackets the forward pass, its "comm_time" stopwatch backward + gradient exchange +
    optimizer step; data_parallel_train.py:111-124).  The split is the measured forward share of the step
    (``probe_step_regions``); without a measurement everything is reported as compute and the source says so."""
    step = regions.get("step_ms", 0.0) or (regions.get("fwd_ms", 0.0) + regions.get("bwd_ms", 0.0) +
                                            regions.get("optimizer_ms", 0.0) + regions.get("exposed_comm_ms", 0.0))
    fwd = regions.get("fwd_ms", 0.0)
    if step > 0 and fwd > 0:
        f = min(max(fwd / step, 0.0), 1.0)
        return dev_s * f, dev_s * (1.0 - f), "device-timed regions"
    return dev_s, 0.0, "unmeasured (region probe off)"


def probe_step_regions(dev, state: List[torch.Tensor], reducers, graphed: "GraphedStep", fwd: Callable, comm: Callable,
                       opt: Callable, restore: Callable, images, labels, iters: int = 10) -> Dict[str, float]:
    """{fwd_ms, bwd_ms, allreduce_ms, optimizer_ms, exposed_comm_ms, step_ms} of one training step.

    The reference brackets regions with ``time.time()`` (data_parallel_train.py:103-124, with the Q8 caveat that its
    "comm_time" is backward+optimizer); here forward and backward (incl. the side-stream wgrads and every collective
    fused into them) are replayed as two separate CUDA graphs and the gradient all-reduces (``comm``) / the optimizer
    pass (``opt``) are launched back to back, every region bracketed by CUDA events and averaged over ``iters``;
    ``exposed_comm_ms`` is what the real (single-graph, overlapped) step costs beyond forward+backward+optimizer.
    ``state`` is snapshotted before and restored after, so the probe does not perturb the run.  Every rank of the
    job must call it together (the regions contain collectives)."""
    dev = torch.device(dev)
    cuda = dev.type == "cuda"
    snap = [t.clone() for t in state]
    reducers = [r for r in reducers if r is not None]
    for r in reducers:
        r.enabled = False

    def clock():
        if cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e
        return time.perf_counter()

    def span(a, b) -> float:
        return a.elapsed_time(b) if cuda else (b - a) * 1e3

    out = {k: 0.0 for k in ("fwd_ms", "bwd_ms", "allreduce_ms", "optimizer_ms", "exposed_comm_ms", "step_ms")}
    mb = None
    is_graphed = cuda and graphed is not None and graphed.graph is not None and tuple(images.shape) == graphed.shape
    try:
        if is_graphed:
            from ..parallel.pp import GraphedMicroBatch
            mb = GraphedMicroBatch(fwd, None, None, dev, True, True, label_shape=tuple(labels.shape), image_like=images)
            torch.cuda.synchronize()
            mb.capture()
            ops.step_end()
            run_f = lambda: mb.run_fwd(images, labels)      # noqa: E731
            run_b = lambda: mb.run_bwd(None)                # noqa: E731
        else:
            holder = {}

            def run_f():
                holder["loss"] = fwd(images, labels)[0]

            def run_b():
                ops.backward(holder.pop("loss"))
                ops.join_side()
                ops.step_end()
        marks = []
        for it in range(iters + 2):
            row = [clock()]
            run_f(); row.append(clock())
            run_b(); row.append(clock())
            comm(); row.append(clock())
            opt(); row.append(clock())
            if it >= 2:
                marks.append(row)
        step_marks = None
        if is_graphed:
            step_marks = [clock()]
            for _ in range(iters):
                graphed.graph.replay()
            step_marks.append(clock())
        if cuda:
            torch.cuda.synchronize()
        n = float(len(marks))
        for i, k in enumerate(("fwd_ms", "bwd_ms", "allreduce_ms", "optimizer_ms")):
            out[k] = sum(span(r[i], r[i + 1]) for r in marks) / n
        if step_marks is not None:
            out["step_ms"] = span(step_marks[0], step_marks[1]) / iters
            out["exposed_comm_ms"] = max(0.0, out["step_ms"] - out["fwd_ms"] - out["bwd_ms"] - out["optimizer_ms"])
        else:
            out["step_ms"] = sum(span(r[0], r[4]) for r in marks) / n
            out["exposed_comm_ms"] = out["allreduce_ms"]        # nothing overlaps on the eager / CPU path
    finally:
        if cuda:
            torch.cuda.synchronize()
        mb = None
        for t, s_ in zip(state, snap):
            t.copy_(s_)
        restore()
        for r in reducers:
            r.enabled = True
            r.begin_step()
    return out


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
