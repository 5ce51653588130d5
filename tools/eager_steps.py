"""A few eager (no CUDA graph) training steps of the flagship engine — the target for
``ncu --metrics gpu__time_duration.sum`` kernel listings (isolated per-kernel durations).
Usage: python tools/eager_steps.py [n_steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops  # noqa: E402
from horizonml_b200.config import TrainConfig  # noqa: E402
from horizonml_b200.trainers.common import Runtime  # noqa: E402
from horizonml_b200.trainers.dp import DPEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
ops.set_backend("native")
cfg = TrainConfig(batch_size=64, device="cuda", dtype="bf16", backend="native", quiet=True, cuda_graph=False)
eng = DPEngine(cfg, Runtime(0, 1, dev, torch.bfloat16, "native", "none"))
g = torch.Generator().manual_seed(0)
x = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(dev)
y = torch.randint(0, 10, (64,), generator=g).to(dev)
for _ in range(n):
    eng.step(x, y)
torch.cuda.synchronize()
print("eager steps done", n)
