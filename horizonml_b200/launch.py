"""Process launcher: spawn, rendezvous, watchdog, failure propagation, result merge.

Reference behaviour (data_parallel_train.py:233-291 and twins): ``mp.set_start_method('spawn')``,
one ``mp.Process`` per rank on ``localhost:<free port>``, join with a shared deadline
``max(120, 120·N/1000)`` s (TP: 400), ``terminate()`` stragglers, merge per-rank CSVs.  Its workers
swallow every exception and exit 0 (Q13) and its exit handshake hangs until the watchdog fires
(Q7), so ``total_training_time`` measures the watchdog, not training.

This launcher keeps the *shape* (spawned local ranks, free port, watchdog, merged CSV, same stdout
lines) and fixes the semantics: clean shutdown, true wall time, non-zero exit + error file on
failure, prompt tear-down of the whole job when one rank dies (also exercised by the
``--inject_fault`` hook), per-rank heartbeat files.  Under ``torchrun`` (RANK in env) no processes
are spawned: the current process is one rank.
"""
from __future__ import annotations

import json
import os
import socket
import sys
import time
import traceback
from typing import Callable, Optional

import torch
import torch.multiprocessing as mp

from .config import TrainConfig
from .metrics import merge_worker_csvs, write_summary

STRATEGY_TITLE = {"data": "Data-parallel", "layer": "Model-parallel", "tensor": "Tensor-parallel"}


def find_free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def default_watchdog_s(strategy: str, sample_size: int) -> float:
    base = 400.0 if strategy == "tensor" else 120.0     # tensor_…:346 vs data_…:250
    return max(base, base * sample_size / 1000.0)


def resolve_device(cfg: TrainConfig) -> str:
    if cfg.device != "auto":
        return cfg.device
    if torch.cuda.is_available() and torch.cuda.device_count() >= cfg.world_size:
        return "cuda"
    return "cpu"


def _worker_entry(rank: int, cfg_json: str, port: int, device: str, train_fn_path: str, err_dir: str):
    cfg = TrainConfig.from_json(cfg_json)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    os.environ["LOCAL_RANK"] = str(rank)
    os.environ["WORLD_SIZE"] = str(cfg.world_size)
    try:
        mod_name, fn_name = train_fn_path.rsplit(":", 1)
        import importlib
        fn = getattr(importlib.import_module(mod_name), fn_name)
        fn(rank, cfg.world_size, cfg, device)
        if not cfg.quiet:
            print(f"Worker {rank} completed successfully", flush=True)
    except BaseException as e:  # noqa: BLE001 - we re-raise after recording
        os.makedirs(err_dir, exist_ok=True)
        with open(os.path.join(err_dir, f"error_rank{rank}.txt"), "w") as fh:
            fh.write("".join(traceback.format_exception(type(e), e, e.__traceback__)))
        print(f"Error in worker {rank}: {e}", file=sys.stderr, flush=True)
        # do not hang peers in a collective: leave hard, the launcher tears the job down
        os._exit(1)


def run_strategy(cfg: TrainConfig, train_fn_path: str):
    """Spawn ``cfg.world_size`` ranks of ``train_fn_path`` ("module:function"), supervise them and
    return the combined DataFrame (or None on failure) — the reference's ``run_*`` contract."""
    logs_dir = cfg.resolved_logs_dir()
    os.makedirs(logs_dir, exist_ok=True)
    for f in os.listdir(logs_dir):
        if f.startswith("error_rank"):
            os.remove(os.path.join(logs_dir, f))
    device = resolve_device(cfg)
    start_time = time.time()

    if "RANK" in os.environ and "WORLD_SIZE" in os.environ and os.environ.get("HZ_SPAWNED") != "1":
        # launched by torchrun: this process IS a rank
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        cfg = cfg.replace(world_size=world)
        mod_name, fn_name = train_fn_path.rsplit(":", 1)
        import importlib
        getattr(importlib.import_module(mod_name), fn_name)(rank, world, cfg, device)
        total = time.time() - start_time
        if rank == 0:
            return merge_worker_csvs(logs_dir, world, cfg.sample_size, total)
        return None

    ctx = mp.get_context("spawn")
    port = find_free_port()
    timeout = cfg.watchdog_s if cfg.watchdog_s > 0 else default_watchdog_s(cfg.strategy, cfg.sample_size)
    if not cfg.quiet:
        print(f"Using port {port} for distributed communication")
        print(f"Using timeout of {timeout} seconds for sample size {cfg.sample_size}")
    env_flag = os.environ.get("HZ_SPAWNED")
    os.environ["HZ_SPAWNED"] = "1"
    procs = []
    try:
        for r in range(cfg.world_size):
            p = ctx.Process(target=_worker_entry,
                            args=(r, cfg.to_json(), port, device, train_fn_path, logs_dir))
            p.start()
            procs.append(p)
    finally:
        if env_flag is None:
            os.environ.pop("HZ_SPAWNED", None)
        else:
            os.environ["HZ_SPAWNED"] = env_flag

    deadline = start_time + timeout
    failed, timed_out = False, False
    while True:
        alive = [p for p in procs if p.is_alive()]
        if any((p.exitcode not in (None, 0)) for p in procs):
            failed = True
            break
        if not alive:
            break
        if time.time() > deadline:
            timed_out = True
            break
        time.sleep(0.05)
    if failed or timed_out:
        why = "a worker failed" if failed else f"watchdog timeout after {timeout:.0f}s"
        print(f"Tearing down job: {why}", file=sys.stderr, flush=True)
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(5)
            if p.is_alive():
                p.kill()
    for p in procs:
        p.join()
    total = time.time() - start_time
    ok = (not failed) and (not timed_out) and all(p.exitcode == 0 for p in procs)
    if not cfg.quiet:
        print(f"{STRATEGY_TITLE[cfg.strategy]} training completed in {total:.2f} seconds")
    combined = merge_worker_csvs(logs_dir, cfg.world_size, cfg.sample_size, total)
    write_summary(logs_dir, f"launch_summary_{cfg.sample_size}.json", {
        "strategy": cfg.strategy, "world_size": cfg.world_size, "device": device, "ok": ok,
        "failed": failed, "timed_out": timed_out, "total_training_time": total,
        "exit_codes": [p.exitcode for p in procs], "config": json.loads(cfg.to_json())})
    if not ok:
        if combined is not None and not cfg.quiet:
            print("Warning: job did not finish cleanly; returning partial results", file=sys.stderr)
        raise_on = os.environ.get("HZ_RAISE_ON_FAILURE", "0") == "1"
        if raise_on:
            raise RuntimeError(f"{cfg.strategy}-parallel job failed (see {logs_dir}/error_rank*.txt)")
        return None
    return combined


def init_distributed(rank: int, world_size: int, device: str, comm: str = "auto") -> str:
    """Rendezvous (reference: setup_distributed, data_parallel_train.py:28-40). Returns backend."""
    import datetime
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    backend = comm if comm != "auto" else ("nccl" if device == "cuda" else "gloo")
    # compute, wgrad side stream, comm stream, one stream per pipeline channel plus NCCL's own: with the default 8
    # hardware queues two of them can share a queue, and a parked receive would then hold up unrelated kernels
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    if os.environ["MASTER_ADDR"] in ("127.0.0.1", "localhost"):
        # one node: gloo would otherwise derive its interface from the host name, which need not resolve in a container
        # (every pair connection then times out); NCCL's bootstrap socket has the same habit
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    kw = {}
    if device == "cuda":
        local = int(os.environ.get("LOCAL_RANK", rank)) % max(torch.cuda.device_count(), 1)
        torch.cuda.set_device(local)
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend, rank=rank, world_size=world_size,
                            timeout=datetime.timedelta(seconds=300), **kw)
    print(f"Process {rank} initialized in a world of {world_size} workers on port "
          f"{os.environ['MASTER_PORT']}", flush=True)
    dist.barrier()
    return backend


def shutdown_distributed() -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        try:
            dist.barrier()
        finally:
            dist.destroy_process_group()
