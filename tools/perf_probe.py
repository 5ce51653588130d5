"""Device-timed measurements of the code written after round 2's last GPU session, one section per process so that a
kernel that traps cannot take the other sections' numbers with it.  Every result is printed at once as a line
``HZPERF <tag> <json>`` (tests/test_gpu_perf_report.py republishes them in the test log, tools/late_suite.sh collects them).

    python tools/perf_probe.py steps      # ResNet-18 / MobileNetV2 training step through the DP engine (native, + PyTorch ops)
    python tools/perf_probe.py handoff    # the same steps with the BatchNorm-backward sums taken in dgrad / pool-backward kernels
    python tools/perf_probe.py conv       # batch-4096 convolutions: one-tile-per-CTA kernel, cuDNN, then the persistent kernels
    python tools/perf_probe.py bench      # bench.py --gpus 1 on the default path and with HZ_BN_BWD_IN_DGRAD=1
    python tools/perf_probe.py bigbatch   # ResNet-18 step at batch 2048: default kernels, PyTorch ops, persistent kernels (auto)

CUDA events after warm-up, synchronised on both sides, a 256 MiB L2-flush write between timed launches."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from horizonml_b200 import ops  # noqa: E402

DEV = "cuda:0"
# HZ_PROBE_DRYRUN=1 (tests/test_cpu_round2.py::test_perf_probe_dry_run): the same code paths on CPU tensors with the extension
# replaced by the test shim — tiny shapes, host clock instead of CUDA events; checks this script, not the kernels
DRYRUN = os.environ.get("HZ_PROBE_DRYRUN", "0") == "1"
SCALE = 1
if DRYRUN:
    DEV, SCALE = "cpu", 64
    os.environ["HZ_GPU_TESTS_DRYRUN"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest as _conftest
    _conftest._install_dryrun_shim()


class _HostEvent:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        import time
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


def Event():
    return _HostEvent() if DRYRUN else torch.cuda.Event(enable_timing=True)


def report(tag, payload):
    print("HZPERF " + tag + " " + json.dumps(payload), flush=True)


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def timed(fn, flush, iters=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        a, b = Event(), Event()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3          # us


def steps(variants, batch=64, persist=0, timed_steps=30):
    import horizonml_b200.models.resnet as R
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.dp import DPEngine
    flush = torch.empty((256 << 20) // (SCALE * SCALE), dtype=torch.uint8, device=DEV)
    batch = max(8, batch // SCALE) if DRYRUN else batch
    g = torch.Generator().manual_seed(0)
    xs = torch.randint(0, 256, (batch, 32, 32, 3), dtype=torch.uint8, generator=g).to(DEV)
    ys = torch.randint(0, 10, (batch,), generator=g).to(DEV)
    for model, be, hand_off in variants:
        try:
            ops.set_backend(be)          # "torch": the same engine on PyTorch ops (cuDNN / ATen kernels) for scale
            R._BN_BWD_IN_DGRAD = hand_off
            if be == "native":
                nb.C.conv_set_persist(persist)      # baked into the captured graph: set before the engine's first step
            cfg = TrainConfig(strategy="data", world_size=1, batch_size=batch, device="cpu" if DRYRUN else "cuda", dtype="bf16",
                              backend=be, model=model, quiet=True, cuda_graph=not DRYRUN)
            eng = DPEngine(cfg, Runtime(0, 1, torch.device(DEV), torch.bfloat16, be, "none"))
            launches = None
            for i in range(6):
                before = sum(nb.LAUNCHES.values())
                eng.step(xs, ys)
                if i == 1:
                    launches = sum(nb.LAUNCHES.values()) - before        # (an eager warm-up step: python-side launches)
            torch.cuda.synchronize()
            K = 2 if DRYRUN else timed_steps
            evs = [(Event(), Event()) for _ in range(K)]
            for a, b in evs:
                flush.fill_(1)
                a.record(); eng.step(xs, ys); b.record()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in evs) / K
            st = eng.stats.buf.float().cpu()           # [sum of losses, correct, samples, ...] over all steps run above
            report("step", {"model": model, "backend": be, "bn_sums_in_dgrad": hand_off, "batch": batch,
                            "conv_persist_mode": persist, "ms_per_step": round(ms, 4),
                            "images_per_s": round(batch / ms * 1e3), "launches_per_step": launches,
                            "mean_loss": round(float(st[0] / max(float(st[5]), 1.0)), 4),
                            "train_acc": round(float(st[1] / max(float(st[2]), 1.0)), 4),
                            "graph": eng._graphed.graph is not None, "fallbacks": dict(nb.FALLBACKS)})
            eng._graphed.graph = None
        except Exception as e:  # noqa: BLE001
            report("step", {"model": model, "backend": be, "bn_sums_in_dgrad": hand_off, "batch": batch,
                            "conv_persist_mode": persist, "error": repr(e)[:300]})
        finally:
            R._BN_BWD_IN_DGRAD = False
            if be == "native":
                nb.C.conv_set_persist(0)


def conv():
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.ops import torch_backend as tb
    ops.set_backend("native")
    peak = 1433.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:  # noqa: BLE001
        pass
    flush = torch.empty((256 << 20) // (SCALE * SCALE), dtype=torch.uint8, device=DEV)
    B = 4096 // SCALE
    data = []
    for name, cin, h, cout in (("layer1", 64, 8, 64), ("layer2", 128, 4, 128), ("layer3", 256, 2, 256)):
        g = torch.Generator().manual_seed(1)
        x = cl((torch.randn(B, cin, h, h, generator=g) * 0.5).to(DEV).bfloat16())
        w = cl((torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(DEV).bfloat16())
        dy = cl((torch.randn(B, cout, h, h, generator=g) * 0.5).to(DEV).bfloat16())
        data.append((name, x, w, dy, 2.0 * B * h * h * cout * cin * 9))

    def kernels(mode, tag):
        for name, x, w, dy, flops in data:
            row = {"layer": name, "batch": B, "kernel": tag, "gflop": round(flops / 1e9, 1)}
            try:
                nb.C.conv_set_persist(mode)
                if mode != 0:                      # numerics next to the timing: same operands through the default kernels
                    y1, s1 = nb.conv_fwd(x, w, 1, 1, True)
                    d1 = nb.conv_dgrad(dy, w, x.shape, 1, 1)
                    nb.C.conv_set_persist(0)
                    y0, s0 = nb.conv_fwd(x, w, 1, 1, True)
                    d0 = nb.conv_dgrad(dy, w, x.shape, 1, 1)
                    nb.C.conv_set_persist(mode)
                    rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))   # noqa: E731
                    row.update(fwd_rel_err_vs_default=round(rel(y1, y0), 6), dgrad_rel_err_vs_default=round(rel(d1, d0), 6),
                               bn_sums_rel_err_vs_default=round(rel(s1.view(-1), s0.view(-1)), 6))
                tf = timed(lambda: nb.conv_fwd(x, w, 1, 1, True), flush)
                td = timed(lambda: nb.conv_dgrad(dy, w, x.shape, 1, 1), flush)
                row.update(fwd_us=round(tf, 1), fwd_tflops=round(flops / tf / 1e6, 1), dgrad_us=round(td, 1),
                           dgrad_tflops=round(flops / td / 1e6, 1), fwd_frac_of_peak=round(flops / tf / 1e6 / peak, 3),
                           peak_tflops=peak)
            except Exception as e:  # noqa: BLE001
                row["error"] = repr(e)[:200]
            finally:
                nb.C.conv_set_persist(0)
            report("conv", row)
    kernels(0, "one_tile_per_cta")
    for name, x, w, dy, flops in data:
        row = {"layer": name, "batch": B, "kernel": "cudnn"}
        try:
            tf = timed(lambda: tb.conv_fwd(x, w, 1, 1, False), flush)
            td = timed(lambda: tb.conv_dgrad(dy, w, x.shape, 1, 1), flush)
            row.update(fwd_us=round(tf, 1), fwd_tflops=round(flops / tf / 1e6, 1), dgrad_us=round(td, 1),
                       dgrad_tflops=round(flops / td / 1e6, 1))
        except Exception as e:  # noqa: BLE001
            row["error"] = repr(e)[:200]
        report("conv", row)
    kernels(2, "persistent_n64")           # (the uncertain kernels last: everything above is already printed)
    kernels(1, "persistent_wide")


def bench_ab():
    """bench.py itself (its timing rules: warm-up, L2 flush, device events, clock sampling, e2e arm) on the default path
    and with the BatchNorm hand-offs: the A/B in the judged harness."""
    import subprocess
    for flag in ("0", "1"):
        row = {"HZ_BN_BWD_IN_DGRAD": int(flag)}
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "100", "--warmup", "10"],
                               cwd=ROOT, env=dict(os.environ, HZ_BN_BWD_IN_DGRAD=flag), capture_output=True, text=True, timeout=100)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{") and '"metric"' in ln]
            if r.returncode != 0 or not line:
                row["error"] = (r.stderr or r.stdout)[-300:].replace("\n", " | ")
            else:
                d = json.loads(line[-1])
                row.update(images_per_s=d.get("value"), ms_per_step=d.get("ms_per_step"), gpu_launches=d.get("gpu_launches"),
                           e2e_images_per_s=(d.get("e2e") or {}).get("value"), clocks=d.get("clocks"), steps=d.get("steps"))
        except Exception as e:  # noqa: BLE001
            row["error"] = repr(e)[:300]
        report("bench", row)


if __name__ == "__main__":
    section = sys.argv[1] if len(sys.argv) > 1 else "steps"
    if section == "steps":
        steps([("resnet18", "native", False), ("mobilenet", "native", False), ("mobilenet", "torch", False)])
    elif section == "handoff":
        steps([("resnet18", "native", True), ("mobilenet", "native", True)])
    elif section == "conv":
        conv()
    elif section == "bench":
        bench_ab()
    elif section == "bigbatch":
        # the throughput regime (many waves of tiles per conv): one-tile-per-CTA kernels, then the persistent kernels
        # where the dispatcher's wave rule picks them (mode -1 = auto), then the same on PyTorch ops for scale
        steps([("resnet18", "native", False)], batch=2048, persist=0, timed_steps=10)
        steps([("resnet18", "torch", False)], batch=2048, timed_steps=10)
        steps([("resnet18", "native", False)], batch=2048, persist=-1, timed_steps=10)
    else:
        raise SystemExit(f"unknown section {section}")
