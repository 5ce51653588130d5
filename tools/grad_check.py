"""Per-parameter gradient agreement of one ResNet-18 training step (B=64):
native bf16 kernels and the PyTorch bf16 oracle, each against an fp32 PyTorch reference."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops  # noqa: E402
from horizonml_b200.models.flat import FlatParams  # noqa: E402
from horizonml_b200.models.resnet import resnet18  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
images = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(dev)
labels = torch.randint(0, 10, (64,), generator=g).to(dev)


def run(backend, dtype):
    ops.set_backend(backend)
    model = resnet18(10, seed=0).to(dev).train()
    flat = FlatParams(list(model.named_parameters()), dev, dtype)
    x = ops.stem_prepare(images.permute(0, 3, 1, 2), dtype=dtype)
    flat.begin_step()
    loss, correct = model.forward_loss(x, labels)
    loss.backward()
    ops.join_side()
    torch.cuda.synchronize()
    return loss.item(), {n: p.main_grad.detach().float().clone() for n, p in zip(flat.names, flat.params)}


ops.enable_side_stream(True)
l32, g32 = run("torch", torch.float32)
lt, gt = run("torch", torch.bfloat16)
ln, gn = run("native", torch.bfloat16)
rows = []
for n in g32:
    a = g32[n].flatten()
    cn = torch.nn.functional.cosine_similarity(a, gn[n].flatten(), dim=0).item()
    ct = torch.nn.functional.cosine_similarity(a, gt[n].flatten(), dim=0).item()
    rows.append({"param": n, "cos_native_vs_fp32": cn, "cos_torchbf16_vs_fp32": ct,
                 "norm_fp32": a.norm().item(), "norm_native": gn[n].norm().item()})
rows.sort(key=lambda r: r["cos_native_vs_fp32"])
allv = lambda d: torch.cat([v.flatten() for v in d.values()])  # noqa: E731
summary = {"loss_fp32": l32, "loss_torch_bf16": lt, "loss_native_bf16": ln,
           "cos_all_native_vs_fp32": torch.nn.functional.cosine_similarity(allv(g32), allv(gn), dim=0).item(),
           "cos_all_torchbf16_vs_fp32": torch.nn.functional.cosine_similarity(allv(g32), allv(gt), dim=0).item(),
           "cos_all_native_vs_torchbf16": torch.nn.functional.cosine_similarity(allv(gt), allv(gn), dim=0).item()}
print(json.dumps(summary))
for r in rows[:12]:
    print(json.dumps(r))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/grad_check.json"
json.dump({"summary": summary, "rows": rows}, open(out, "w"), indent=1)
