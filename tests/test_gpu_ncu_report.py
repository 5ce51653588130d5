"""Nsight Compute over one training step, inside the GPU test run: `ncu --profile-from-start off` around
tools/ncu_step_probe.py (one eager ResNet-18 step after warm-up), a handful of raw metrics per launch, aggregated per
kernel and published as ``HZPERF ncu`` lines in pytest's warnings summary — the profile of the tree that is being judged
(the committed capture under profiles/ is round 1's).  Numbers taken under the profiler are shares and utilisations, never
benchmark values.  `late` (order 10: right after the numerics tests — it profiles the default, hardware-verified path, so
it does not depend on any of the new kernels — and before the timing sections), time-boxed: a run that cannot attach or
finish is a skip."""
import collections
import csv
import io
import json
import os
import shutil
import signal
import subprocess
import sys
import warnings

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=10)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum"]


def _short(name: str) -> str:
    name = name.split("(")[0]
    for pre in ("void hz::", "hz::", "void "):
        if name.startswith(pre):
            name = name[len(pre):]
    return name[:60]


def test_ncu_profile_of_one_training_step():
    exe = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(exe):
        pytest.skip("ncu not installed")
    cmd = [exe, "--profile-from-start", "off", "--clock-control", "none", "--metrics", ",".join(METRICS), "--csv", "--page", "raw",
           sys.executable, os.path.join(ROOT, "tools", "ncu_step_probe.py")]
    env = dict(os.environ, CUDA_MODULE_LOADING="LAZY")
    proc = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                            start_new_session=True)
    try:
        out, err = proc.communicate(timeout=110)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except OSError:
            pass
        proc.communicate()
        pytest.skip("ncu run did not finish within 110 s")
    start = out.find('"ID"')
    if start < 0:
        pytest.skip("ncu produced no CSV (permissions?): " + (err or out)[-300:].replace("\n", " | "))
    rows = list(csv.reader(io.StringIO(out[start:])))
    header, rows = rows[0], [r for r in rows[2:] if len(r) == len(rows[0])]      # (row 1 = units)
    col = {h: i for i, h in enumerate(header)}
    if "Kernel Name" not in col or METRICS[0] not in col:
        pytest.skip("unexpected ncu CSV layout: " + ",".join(header)[:300])

    def num(r, m):
        try:
            return float(r[col[m]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")
    agg = collections.OrderedDict()
    for r in rows:
        k = _short(r[col["Kernel Name"]])
        a = agg.setdefault(k, {"launches": 0, "ns": 0.0, "sm": 0.0, "dram": 0.0, "occ": 0.0, "regs": 0, "bytes": 0.0})
        t = num(r, METRICS[0])
        a["launches"] += 1
        a["ns"] += t
        a["sm"] += num(r, METRICS[1]) * t                      # time-weighted utilisations
        a["dram"] += num(r, METRICS[2]) * t
        a["occ"] += num(r, METRICS[3]) * t
        a["regs"] = max(a["regs"], int(num(r, METRICS[4]) or 0))
        a["bytes"] += num(r, METRICS[5]) + num(r, METRICS[6])
    total = sum(a["ns"] for a in agg.values()) or 1.0
    warnings.warn("HZPERF ncu_total " + json.dumps({"kernels": len(rows), "distinct": len(agg), "sum_of_durations_us": round(total / 1e3, 1),
                                                   "note": "eager step, serialised under the profiler: shares, not a step time"}))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])[:14]:
        t = a["ns"] or 1.0
        warnings.warn("HZPERF ncu " + json.dumps({
            "kernel": k, "launches": a["launches"], "us": round(a["ns"] / 1e3, 1), "share": round(a["ns"] / total, 3),
            "sm_pct": round(a["sm"] / t, 1), "dram_pct": round(a["dram"] / t, 1), "warps_active_pct": round(a["occ"] / t, 1),
            "regs": a["regs"], "dram_MB": round(a["bytes"] / 1e6, 2)}))
    assert len(rows) > 50
