"""Per-epoch metrics, CSV schema and summaries.

Reference schema (exact column order, data_parallel_train.py:161-173; layer/tensor insert
``avg_bandwidth`` before ``grad_divergence``: layer_…:289-302, tensor_…:266-279):

    epoch, loss, accuracy, epoch_time, avg_step_time, compute_time, comm_time, idle_time,
    avg_cpu, avg_memory, [avg_bandwidth,] grad_divergence

Reference conventions kept for drop-in analysis: compute/comm/idle are cumulative
over epochs, avg_* reset per epoch, ``avg_bandwidth`` is bytes per step.  Extended columns
(device-timed, appended after the reference columns) add what the reference could not measure.
"""
from __future__ import annotations

import json
import os
import time
from typing import Dict, List, Optional

import pandas as pd

try:
    import psutil
except Exception:  # pragma: no cover
    psutil = None

REF_COLUMNS_DP = ["epoch", "loss", "accuracy", "epoch_time", "avg_step_time", "compute_time",
                  "comm_time", "idle_time", "avg_cpu", "avg_memory", "grad_divergence"]
REF_COLUMNS_BW = REF_COLUMNS_DP[:-1] + ["avg_bandwidth", "grad_divergence"]
EXT_COLUMNS = ["images_per_sec", "fwd_ms", "bwd_ms", "allreduce_ms", "exposed_comm_ms", "p2p_ms",
               "optimizer_ms", "nvlink_GBps", "gpu_mem_MB", "steps", "split_source"]


def ref_columns(strategy: str) -> List[str]:
    return list(REF_COLUMNS_DP if strategy == "data" else REF_COLUMNS_BW)


class HostSampler:
    """psutil CPU% / RSS sampler (data_parallel_train.py:105-106)."""

    def __init__(self):
        self.proc = psutil.Process(os.getpid()) if psutil else None
        self.cpu: List[float] = []
        self.mem: List[float] = []

    def sample(self) -> None:
        if self.proc is None:
            return
        self.cpu.append(psutil.cpu_percent(interval=None))
        self.mem.append(self.proc.memory_info().rss / 1024 / 1024)

    def drain(self):
        c = sum(self.cpu) / len(self.cpu) if self.cpu else 0.0
        m = sum(self.mem) / len(self.mem) if self.mem else 0.0
        self.cpu, self.mem = [], []
        return c, m


class EpochRecorder:
    """Accumulates the reference's per-epoch record and rewrites the per-rank CSV every epoch."""

    def __init__(self, strategy: str, rank: int, logs_dir: str, sample_size: int):
        self.strategy, self.rank, self.logs_dir = strategy, rank, logs_dir
        self.sample_size = sample_size
        os.makedirs(logs_dir, exist_ok=True)
        self.rows: List[Dict] = []
        self.total_compute = 0.0     # cumulative across epochs (reference convention)
        self.total_comm = 0.0
        self.total_idle = 0.0
        self.grad_divs: List[float] = []
        self.host = HostSampler()

    @property
    def path(self) -> str:
        # Q11 fix: name by requested sample_size (what the launcher merges on)
        return os.path.join(self.logs_dir, f"worker_{self.rank}_samples_{self.sample_size}.csv")

    def end_epoch(self, epoch: int, loss: float, accuracy: float, epoch_time: float,
                  step_times: List[float], avg_bandwidth: Optional[float] = None,
                  ext: Optional[Dict] = None) -> Dict:
        cpu, mem = self.host.drain()
        row = {
            "epoch": epoch, "loss": loss, "accuracy": accuracy, "epoch_time": epoch_time,
            "avg_step_time": sum(step_times) / len(step_times) if step_times else 0,
            "compute_time": self.total_compute, "comm_time": self.total_comm,
            "idle_time": self.total_idle, "avg_cpu": cpu, "avg_memory": mem,
        }
        if self.strategy != "data":
            row["avg_bandwidth"] = avg_bandwidth if avg_bandwidth is not None else 0
        row["grad_divergence"] = (sum(self.grad_divs) / len(self.grad_divs)) if self.grad_divs else 0
        for k in EXT_COLUMNS:
            row[k] = (ext or {}).get(k, 0)
        self.rows.append(row)
        self.write()
        return row

    def frame(self) -> pd.DataFrame:
        cols = ref_columns(self.strategy) + EXT_COLUMNS
        return pd.DataFrame(self.rows, columns=cols)

    def write(self) -> None:
        self.frame().to_csv(self.path, index=False)


def merge_worker_csvs(logs_dir: str, world_size: int, sample_size: int,
                      total_training_time: float) -> Optional[pd.DataFrame]:
    """Launcher-side merge → combined_results_{N}.csv (data_parallel_train.py:276-291)."""
    frames = []
    for rank in range(world_size):
        f = os.path.join(logs_dir, f"worker_{rank}_samples_{sample_size}.csv")
        if os.path.exists(f):
            try:
                df = pd.read_csv(f)
            except Exception:
                continue
            df["worker"] = rank
            df["total_training_time"] = total_training_time
            frames.append(df)
    if not frames:
        return None
    combined = pd.concat(frames, ignore_index=True)
    combined.to_csv(os.path.join(logs_dir, f"combined_results_{sample_size}.csv"), index=False)
    return combined


def write_summary(logs_dir: str, name: str, payload: Dict) -> str:
    os.makedirs(logs_dir, exist_ok=True)
    path = os.path.join(logs_dir, name)
    with open(path, "w") as fh:
        json.dump(payload, fh, indent=1, sort_keys=True, default=str)
    return path


class Stopwatch:
    """Host stopwatch with the reference's region names (time.time() pairs)."""

    def __init__(self):
        self.t0 = time.time()

    def lap(self) -> float:
        t = time.time()
        d, self.t0 = t - self.t0, t
        return d
