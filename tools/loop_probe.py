"""Where does the trainer loop spend host time? (diagnosis tool)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops
from horizonml_b200.config import TrainConfig
from horizonml_b200.data import BatchLoader, build_dataset
from horizonml_b200.metrics import HostSampler
from horizonml_b200.trainers.common import Runtime
from horizonml_b200.trainers.dp import DPEngine

dev = torch.device("cuda", 0)
ops.set_backend("native")
cfg = TrainConfig(batch_size=64, device="cuda", dtype="bf16", backend="native", quiet=True)
eng = DPEngine(cfg, Runtime(0, 1, dev, torch.bfloat16, "native", "none"))
images, labels = build_dataset(16384, True, "./data", 1)
loader = BatchLoader(images, labels, 64, dev)
hs = HostSampler()

def timed(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); n = fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"{name:40s} {dt*1e3/n:8.3f} ms/iter ({n} iters)", flush=True)

def only_loader():
    n = 0
    for x, y in loader: n += 1
    return n
def loader_step():
    n = 0
    for x, y in loader:
        eng.step(x, y); n += 1
    return n
def loader_step_sample():
    n = 0
    for x, y in loader:
        hs.sample(); eng.step(x, y); n += 1
    return n
x0 = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, device=dev); y0 = torch.randint(0, 10, (64,), device=dev)
def step_only():
    for _ in range(256): eng.step(x0, y0)
    return 256
def sample_only():
    for _ in range(256): hs.sample()
    return 256
for _ in range(10): eng.step(x0, y0)
timed("step only (device-resident batch)", step_only)
timed("host sampler only", sample_only)
timed("loader only", only_loader)
timed("loader + step", loader_step)
timed("loader + step + host sampler", loader_step_sample)

# ---- finer: host time per phase inside the combined loop
import collections
acc = collections.Counter()
it = iter(loader)
n = 0
torch.cuda.synchronize()
while True:
    t0 = time.perf_counter()
    try:
        x, y = next(it)
    except StopIteration:
        break
    t1 = time.perf_counter()
    eng.step(x, y)
    t2 = time.perf_counter()
    acc["next(loader)"] += t1 - t0; acc["eng.step"] += t2 - t1; n += 1
torch.cuda.synchronize()
print({k: round(v * 1e3 / n, 3) for k, v in acc.items()}, "ms/iter host time")
# inside the loader: time the pieces of _stage
import numpy as np
acc2 = collections.Counter()
orig_stage = loader._stage
def stage_probe(bidx, slot):
    t0 = time.perf_counter()
    if loader._copied[slot] is not None:
        loader._copied[slot].synchronize()
    t1 = time.perf_counter()
    t = torch.from_numpy(np.ascontiguousarray(bidx))
    torch.index_select(loader.images, 0, t, out=loader._hx[slot][:len(bidx)])
    torch.index_select(loader.labels, 0, t, out=loader._hy[slot][:len(bidx)])
    t2 = time.perf_counter()
    r = orig_stage(bidx, slot)
    t3 = time.perf_counter()
    acc2["copied.synchronize"] += t1 - t0; acc2["gather"] += t2 - t1; acc2["orig_stage(total again)"] += t3 - t2
    return r
loader._stage = stage_probe
n = 0
for x, y in loader:
    eng.step(x, y); n += 1
torch.cuda.synchronize()
print({k: round(v * 1e3 / n, 3) for k, v in acc2.items()}, "ms/iter inside _stage")
print("pinned?", loader._hx[0].is_pinned(), loader._hx[0][:64].is_pinned())
