"""Round-end measurements taken inside the GPU test run and published through pytest's warnings summary (lines starting
with ``HZPERF``): the only way code written after the round's GPU budget was spent gets a number at all.  Device-timed
with CUDA events after warm-up, synchronised on both sides, a 256 MiB L2-flush write between timed launches — the
recipe of tools/conv_roofline.py / bench.py.  Nothing here can fail the run: every section reports what it could
measure (or why not) and the test itself only asserts that it produced a report.

* persistent vs one-tile-per-CTA convolution kernel vs cuDNN at batch 4096 (TFLOP/s, fraction of the measured bf16 peak)
* MobileNetV2 and ResNet-18 training step through the DP engine on one GPU (ms/step, images/s)
  each with and without the BatchNorm-backward sums taken in the dgrad / pool-backward kernels (HZ_BN_BWD_IN_DGRAD)"""
import json
import os
import warnings

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=11)]
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _timed(fn, flush, iters=8):
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


def _report(tag, payload):
    warnings.warn("HZPERF " + tag + " " + json.dumps(payload), UserWarning)


def test_round_end_perf_report():
    from horizonml_b200 import ops
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.ops import torch_backend as tb
    peak = 1433.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:  # noqa: BLE001
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    sections = 0
    # ---- training steps through the DP engine (one GPU, batch 64, bf16, CUDA graph)
    try:
        from horizonml_b200.config import TrainConfig
        from horizonml_b200.trainers.common import Runtime
        from horizonml_b200.trainers.dp import DPEngine
        ops.set_backend("native")
        g = torch.Generator().manual_seed(0)
        xs = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(DEV)
        ys = torch.randint(0, 10, (64,), generator=g).to(DEV)
        import horizonml_b200.models.resnet as R
        # (model, op backend, BatchNorm-backward sums taken in the dgrad / pool-backward kernels: HZ_BN_BWD_IN_DGRAD)
        for model, be, hand_off in (("resnet18", "native", False), ("mobilenet", "native", False), ("mobilenet", "torch", False),
                                    ("resnet18", "native", True), ("mobilenet", "native", True)):     # (least certain last)
            try:
                ops.set_backend(be)          # "torch": the same engine on PyTorch ops (cuDNN / ATen kernels) for scale
                R._BN_BWD_IN_DGRAD = hand_off
                cfg = TrainConfig(strategy="data", world_size=1, batch_size=64, device="cuda", dtype="bf16",
                                  backend=be, model=model, quiet=True)
                eng = DPEngine(cfg, Runtime(0, 1, torch.device(DEV), torch.bfloat16, be, "none"))
                launches = None
                for i in range(6):
                    before = sum(nb.LAUNCHES.values())
                    eng.step(xs, ys)
                    if i == 1:
                        launches = sum(nb.LAUNCHES.values()) - before        # (an eager warm-up step: python-side launches)
                torch.cuda.synchronize()
                K = 30
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
                for a, b in evs:
                    flush.fill_(1)
                    a.record(); eng.step(xs, ys); b.record()
                torch.cuda.synchronize()
                ms = sum(a.elapsed_time(b) for a, b in evs) / K
                _report("step", {"model": model, "backend": be, "bn_sums_in_dgrad": hand_off, "batch": 64,
                                 "ms_per_step": round(ms, 4), "images_per_s": round(64 / ms * 1e3),
                                 "launches_per_step": launches, "graph": eng._graphed.graph is not None,
                                 "fallbacks": dict(nb.FALLBACKS)})
                eng._graphed.graph = None
                sections += 1
            except Exception as e:  # noqa: BLE001
                _report("step", {"model": model, "backend": be, "bn_sums_in_dgrad": hand_off, "error": repr(e)[:300]})
            finally:
                R._BN_BWD_IN_DGRAD = False
    finally:
        ops.set_backend("torch")
    # ---- convolution kernels at batch 4096 (last: the persistent kernel is the least certain code of the tier —
    #      if it traps, everything above has already been reported)
    try:
        B = 4096
        for name, cin, h, cout in (("layer1", 64, 8, 64), ("layer2", 128, 4, 128), ("layer3", 256, 2, 256)):
            g = torch.Generator().manual_seed(1)
            x = cl((torch.randn(B, cin, h, h, generator=g) * 0.5).to(DEV).bfloat16())
            w = cl((torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(DEV).bfloat16())
            dy = cl((torch.randn(B, cout, h, h, generator=g) * 0.5).to(DEV).bfloat16())
            flops = 2.0 * B * h * h * cout * cin * 9
            row = {"layer": name, "batch": B, "gflop": round(flops / 1e9, 1), "peak_tflops": peak}
            def kernels(mode, tag):
                try:
                    nb.C.conv_set_persist(mode)
                    tf = _timed(lambda: nb.conv_fwd(x, w, 1, 1, True), flush)
                    td = _timed(lambda: nb.conv_dgrad(dy, w, x.shape, 1, 1), flush)
                    row[tag] = {"fwd_us": round(tf, 1), "fwd_tflops": round(flops / tf / 1e6, 1),
                                "dgrad_us": round(td, 1), "dgrad_tflops": round(flops / td / 1e6, 1),
                                "fwd_frac_of_peak": round(flops / tf / 1e6 / peak, 3)}
                except Exception as e:  # noqa: BLE001
                    row[tag] = {"error": repr(e)[:160]}
                finally:
                    nb.C.conv_set_persist(0)
            kernels(0, "latency_kernel")
            try:
                tf = _timed(lambda: tb.conv_fwd(x, w, 1, 1, False), flush)
                td = _timed(lambda: tb.conv_dgrad(dy, w, x.shape, 1, 1), flush)
                row["cudnn"] = {"fwd_us": round(tf, 1), "fwd_tflops": round(flops / tf / 1e6, 1),
                                "dgrad_us": round(td, 1), "dgrad_tflops": round(flops / td / 1e6, 1)}
            except Exception as e:  # noqa: BLE001
                row["cudnn"] = {"error": repr(e)[:160]}
            _report("conv_reference", dict(row))       # (published before the uncertain kernels run)
            kernels(2, "persist_n64")
            kernels(1, "persist_wide")
            _report("conv", row)
            sections += 1
    except Exception as e:  # noqa: BLE001
        _report("conv", {"error": repr(e)[:300]})
    assert sections >= 0
