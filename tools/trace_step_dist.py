"""Multi-rank kernel timeline of the data-parallel step (torch.profiler / CUPTI on CUDA-graph replays, every rank):
which kernel runs when on which stream — where the gradient all-reduces sit relative to backward, what is exposed
after the last backward kernel.  Diagnosis only, never a benchmark number (profiling perturbs the step).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/trace_step_dist.py out_prefix [bench flags]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_dp"
    flags = sys.argv[2:]
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.trainers.common import setup_runtime
    from horizonml_b200.trainers.dp import DPEngine
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    cfg = TrainConfig(strategy="data", world_size=world, batch_size=64, device="cuda", dtype="bf16", backend="native",
                      quiet=True, fused_adam="--fused_adam" in flags,
                      bucket_layout="layers" if "--layers" in flags else ("size" if "--size" in flags else "auto"))
    rt = setup_runtime(rank, world, cfg, "cuda")
    eng = DPEngine(cfg, rt)
    g = torch.Generator().manual_seed(rank)
    x = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(rt.device)
    y = torch.randint(0, 10, (64,), generator=g).to(rt.device)
    for _ in range(10):
        eng.step(x, y)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            eng.step(x, y)
        torch.cuda.synchronize()
    rows = []
    for ev in prof.profiler.kineto_results.events():
        if ev.device_type() == torch.autograd.DeviceType.CUDA and ev.duration_ns() > 0:
            rows.append({"name": ev.name()[:70], "stream": int(ev.device_resource_id()), "start_us": ev.start_ns() / 1e3,
                         "dur_us": ev.duration_ns() / 1e3})
    rows.sort(key=lambda r: r["start_us"])
    n = len(rows) // 5
    step = rows[3 * n:4 * n]
    t0 = step[0]["start_us"]
    for r in step:
        r["start_us"] = round(r["start_us"] - t0, 2)
        r["dur_us"] = round(r["dur_us"], 2)
    streams = sorted({r["stream"] for r in step}, key=lambda s: -sum(1 for r in step if r["stream"] == s))
    main_s = streams[0]
    comm = [r for r in step if "allreduce_kernel" in r["name"]]
    adam = [r for r in step if "adam_kernel" in r["name"]]
    bwd_main = [r for r in step if r["stream"] == main_s and "allreduce" not in r["name"] and "adam" not in r["name"]
                and "stats_update" not in r["name"]]
    end = max(r["start_us"] + r["dur_us"] for r in step)
    last_compute_end = max(r["start_us"] + r["dur_us"] for r in bwd_main) if bwd_main else 0
    summary = {"rank": rank, "world": world, "kernels": len(step), "span_us": round(end, 1),
               "last_fwd_bwd_kernel_end_us": round(last_compute_end, 1),
               "exposed_after_backward_us": round(end - last_compute_end, 1),
               "allreduce": [[r["start_us"], r["dur_us"]] for r in comm], "adam": [[r["start_us"], r["dur_us"]] for r in adam],
               "streams": len(streams), "fused_adam": eng.fused_adam, "buckets": len(eng.flat.buckets)}
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    json.dump({"summary": summary, "step": step}, open(f"{prefix}_rank{rank}.json", "w"))
    if rank == 0:
        print(json.dumps(summary), flush=True)
    eng._graphed.graph = None
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
