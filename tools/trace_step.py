"""Kernel timeline of one (graph-replayed) training step via torch.profiler/CUPTI — diagnosis only,
never a benchmark number.  Usage: python tools/trace_step.py gpurun_out/trace.json [--no_graph]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops  # noqa: E402
from horizonml_b200.config import TrainConfig  # noqa: E402
from horizonml_b200.trainers.common import Runtime  # noqa: E402
from horizonml_b200.trainers.dp import DPEngine  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.json"
use_graph = "--no_graph" not in sys.argv
dev = torch.device("cuda", 0)
ops.set_backend("native")
cfg = TrainConfig(batch_size=64, device="cuda", dtype="bf16", backend="native", quiet=True, cuda_graph=use_graph)
eng = DPEngine(cfg, Runtime(0, 1, dev, torch.bfloat16, "native", "none"))
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(dev)
y = torch.randint(0, 10, (64,), generator=g).to(dev)
for _ in range(8):
    eng.step(x, y)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        eng.step(x, y)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
rows = [{"name": e.name[:80], "start_us": e.time_range.start, "dur_us": e.time_range.end - e.time_range.start} for e in evs]
if rows:
    t0 = rows[0]["start_us"]
    for r in rows:
        r["start_us"] -= t0
# split into the 3 steps by the largest gaps
n = len(rows) // 3
step = rows[n:2 * n]
busy = sum(r["dur_us"] for r in step)
span = (step[-1]["start_us"] + step[-1]["dur_us"] - step[0]["start_us"]) if step else 0
gaps = [step[i + 1]["start_us"] - (step[i]["start_us"] + step[i]["dur_us"]) for i in range(len(step) - 1)]
summary = {"graph": eng._graphed.graph is not None, "kernels_per_step": n, "span_us": span, "busy_us": busy,
           "gap_us_total": sum(g for g in gaps if g > 0), "gap_us_median": sorted(gaps)[len(gaps) // 2] if gaps else 0}
print(json.dumps(summary))
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump({"summary": summary, "step": step}, open(out, "w"), indent=0)
for r in step:
    print(f"{r['start_us'] - step[0]['start_us']:9.1f} {r['dur_us']:7.1f}  {r['name']}")
