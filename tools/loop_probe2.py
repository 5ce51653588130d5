import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops
from horizonml_b200.config import TrainConfig
from horizonml_b200.data import BatchLoader, build_dataset
from horizonml_b200.trainers.common import Runtime
from horizonml_b200.trainers.dp import DPEngine
dev = torch.device("cuda", 0)
ops.set_backend("native")
graph = os.environ.get("NOGRAPH", "0") != "1"
cfg = TrainConfig(batch_size=64, device="cuda", dtype="bf16", backend="native", quiet=True, cuda_graph=graph)
eng = DPEngine(cfg, Runtime(0, 1, dev, torch.bfloat16, "native", "none"))
images, labels = build_dataset(16384, True, "./data", 1)
x0 = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, device=dev); y0 = torch.randint(0, 10, (64,), device=dev)
for _ in range(10): eng.step(x0, y0)
def run(name, fn, n=256):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(n); torch.cuda.synchronize()
    print(f"{name:50s} {(time.perf_counter()-t)*1e3/n:8.3f} ms/iter", flush=True)
run("step only", lambda n: [eng.step(x0, y0) for _ in range(n)])
# A: pinned host batch -> H2D on the SAME stream each step
hx = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8).pin_memory(); hy = torch.randint(0, 10, (64,)).pin_memory()
def a(n):
    for _ in range(n):
        eng.step(hx.to(dev, non_blocking=True), hy.to(dev, non_blocking=True))
run("H2D on compute stream + step", a)
# B: H2D on a side stream with event to compute stream (no consumed back-edge)
cs = torch.cuda.Stream()
dxs = [torch.empty_like(x0) for _ in range(4)]; dys = [torch.empty_like(y0) for _ in range(4)]
def b(n):
    for i in range(n):
        with torch.cuda.stream(cs):
            dxs[i % 4].copy_(hx, non_blocking=True); dys[i % 4].copy_(hy, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs)
        torch.cuda.current_stream().wait_event(ev)
        eng.step(dxs[i % 4], dys[i % 4])
run("H2D on copy stream (+event) + step", b)
# C: like B plus the consumed back-edge (copy stream waits for the step that used the slot)
cons = [None] * 4
def c(n):
    for i in range(n):
        with torch.cuda.stream(cs):
            if cons[i % 4] is not None: cs.wait_event(cons[i % 4])
            dxs[i % 4].copy_(hx, non_blocking=True); dys[i % 4].copy_(hy, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs)
        torch.cuda.current_stream().wait_event(ev)
        eng.step(dxs[i % 4], dys[i % 4])
        e2 = torch.cuda.Event(); e2.record(torch.cuda.current_stream()); cons[i % 4] = e2
run("copy stream + consumed back-edge + step", c)
loader = BatchLoader(images, labels, 64, dev)
def d(n):
    k = 0
    for x, y in loader:
        eng.step(x, y); k += 1
run("BatchLoader + step", d)
