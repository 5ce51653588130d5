"""One ResNet-18 training step (batch 64, bf16, eager — no CUDA graph, so every kernel is a separate launch) between
cudaProfilerStart / cudaProfilerStop, after two warm-up steps: the capture window of `ncu --profile-from-start off`
(tests/test_gpu_ncu_report.py, tools/profile_1gpu.sh).  HZ_BN_BWD_IN_DGRAD etc. are honoured like anywhere else."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops  # noqa: E402
from horizonml_b200.models.flat import FlatAdam, FlatParams  # noqa: E402
from horizonml_b200.models.resnet import resnet18  # noqa: E402

dev = torch.device("cuda", 0)
ops.set_backend("native")
model = resnet18(10, seed=0).to(dev).train()
flat = FlatParams(list(model.named_parameters()), dev, torch.bfloat16)
opt = FlatAdam(flat, lr=1e-3)
g = torch.Generator().manual_seed(0)
images = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(dev)
labels = torch.randint(0, 10, (64,), generator=g).to(dev)


def step():
    x = ops.stem_prepare(images.permute(0, 3, 1, 2), dtype=torch.bfloat16)
    ops.step_begin(dev)
    flat.begin_step()
    loss, _ = model.forward_loss(x, labels)
    ops.backward(loss)
    ops.join_side()
    ops.step_end()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("probe done")
