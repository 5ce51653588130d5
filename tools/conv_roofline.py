"""Throughput of the tcgen05 implicit-GEMM convolution kernels (forward / dgrad / wgrad) at ResNet-18's layer
shapes for growing batch sizes, next to cuDNN (torch.nn.functional / aten) on the same tensors — where the
hand-written kernels sit against the measured bf16 peak once the problem is large enough to be compute-bound.
Device-timed (CUDA events, after warm-up); a 256 MiB L2 flush between timed launches.
With HZ_CONV_PERSIST=1 (or `auto`) the forward / dgrad columns are additionally measured on the persistent kernel
(the plain columns are then the persistent kernel's, `fwd_default_us` / `dgrad_default_us` the one-tile-per-CTA
kernel's; igemm_persist_kernel: one CTA per SM, two TMEM accumulators, epilogue under the next tile's MMAs).
Usage: python tools/conv_roofline.py gpurun_out/conv_roofline.json [batches...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizonml_b200 import ops  # noqa: E402
from horizonml_b200.ops import native_backend as nb  # noqa: E402
from horizonml_b200.ops import torch_backend as tb  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_roofline.json"
batches = [int(a) for a in sys.argv[2:]] or [64, 512, 4096]
dev = torch.device("cuda", 0)
ops.set_backend("native")
PEAK = 1433.5e12
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] * 1e12
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3          # us


# (Cin, H, Cout, R, stride, pad): one representative conv per ResNet-18 stage at 32x32 input
LAYERS = [("layer1.conv", 64, 8, 64, 3, 1, 1), ("layer2.conv", 128, 4, 128, 3, 1, 1),
          ("layer3.conv", 256, 2, 256, 3, 1, 1), ("layer2.0.conv1(s2)", 64, 8, 128, 3, 2, 1)]
rows = []
for B in batches:
    for name, cin, h, cout, r, s, p in LAYERS:
        g = torch.Generator().manual_seed(1)
        x = cl((torch.randn(B, cin, h, h, generator=g) * 0.5).to(dev).bfloat16())
        w = cl((torch.randn(cout, cin, r, r, generator=g) / (cin * r * r) ** 0.5).to(dev).bfloat16())
        ho = (h + 2 * p - r) // s + 1
        dy = cl((torch.randn(B, cout, ho, ho, generator=g) * 0.5).to(dev).bfloat16())
        flops = 2.0 * B * ho * ho * cout * cin * r * r
        gbuf = torch.zeros(cout * r * r * cin, device=dev)
        gv = gbuf.view(cout, r, r, cin).permute(0, 3, 1, 2)
        gref = torch.zeros(cout, cin, r, r, device=dev)
        row = {"batch": B, "layer": name, "shape": [B, cin, h, h, cout, r, s], "gflop": flops / 1e9}
        try:
            before = sum(nb.FALLBACKS.values())
            t = {"fwd": timed(lambda: nb.conv_fwd(x, w, s, p, True)),
                 "dgrad": timed(lambda: nb.conv_dgrad(dy, w, x.shape, s, p)),
                 "wgrad": timed(lambda: nb.conv_wgrad(dy, x, w.shape, s, p, gv, False))}
            row["native_fallbacks"] = sum(nb.FALLBACKS.values()) - before
            if os.environ.get("HZ_CONV_PERSIST", "0")[:1] in ("1", "a"):
                nb.C.conv_set_persist(0)        # the columns above were taken in the requested mode: add the default kernel's
                row["fwd_default_us"] = round(timed(lambda: nb.conv_fwd(x, w, s, p, True)), 2)
                row["dgrad_default_us"] = round(timed(lambda: nb.conv_dgrad(dy, w, x.shape, s, p)), 2)
                nb.C.conv_set_persist(1 if os.environ["HZ_CONV_PERSIST"][:1] == "1" else -1)
            c = {"fwd": timed(lambda: tb.conv_fwd(x, w, s, p, False)),
                 "dgrad": timed(lambda: tb.conv_dgrad(dy, w, x.shape, s, p)),
                 "wgrad": timed(lambda: tb.conv_wgrad(dy, x, w.shape, s, p, gref, False))}
            for k in t:
                row[f"{k}_us"] = round(t[k], 2)
                row[f"{k}_tflops"] = round(flops / t[k] / 1e6, 1)
                row[f"{k}_frac_of_peak"] = round(flops / (t[k] * 1e-6) / PEAK, 4)
                row[f"{k}_cudnn_us"] = round(c[k], 2)
                row[f"{k}_speedup_vs_cudnn"] = round(c[k] / t[k], 2)
        except Exception as e:  # noqa: BLE001
            row["error"] = repr(e)[:200]
        rows.append(row)
        print(json.dumps(row), flush=True)
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump({"peak_bf16_tflops": PEAK / 1e12, "rows": rows}, open(out, "w"), indent=1)
