"""Nsight Compute over one training step, inside the GPU test run: `ncu --profile-from-start off` around
tools/ncu_step_probe.py (one eager ResNet-18 step after warm-up), a handful of raw metrics per launch, aggregated per
kernel and published as ``HZPERF ncu`` lines in pytest's warnings summary — the profile of the tree that is being judged
(the committed capture under profiles/ is round 1's).  Numbers taken under the profiler are shares and utilisations, never
benchmark values.  `late` (order 10: right after the numerics tests — it profiles the default, hardware-verified path, so
it does not depend on any of the new kernels — and before the timing sections), time-boxed: a run that cannot attach or
finish is a skip."""
import json
import os
import shutil
import signal
import subprocess
import sys
import warnings

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=10, limit_s=190)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402  (the aggregation is shared with the offline tool)

METRICS = ncu_summary.BASE_METRICS


TENSOR = ncu_summary.TENSOR                     # tensor-pipe utilisation (optional: older ncu builds may not know the name)


def _ncu(exe, metrics, budget_s):
    """(stdout, stderr) of one profiled probe run, or None if it did not finish in time"""
    cmd = [exe, "--profile-from-start", "off", "--clock-control", "none", "--metrics", ",".join(metrics), "--csv", "--page", "raw",
           sys.executable, os.path.join(ROOT, "tools", "ncu_step_probe.py")]
    proc = subprocess.Popen(cmd, cwd=ROOT, env=dict(os.environ, CUDA_MODULE_LOADING="LAZY"), stdout=subprocess.PIPE,
                            stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        return proc.communicate(timeout=budget_s)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except OSError:
            pass
        proc.communicate()
        return None


def test_ncu_profile_of_one_training_step():
    import time
    exe = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(exe):
        pytest.skip("ncu not installed")
    t0 = time.time()
    res = _ncu(exe, METRICS + [TENSOR], 80)
    if res is not None and res[0].find('"ID"') < 0 and time.time() - t0 < 30:
        res = _ncu(exe, METRICS, 80)           # refused quickly (a metric name this ncu does not know): base list
    if res is None:
        pytest.skip("ncu run did not finish within its time box")
    out, err = res
    start = out.find('"ID"')
    if start < 0:
        pytest.skip("ncu produced no CSV (permissions?): " + (err or out)[-300:].replace("\n", " | "))
    res = ncu_summary.aggregate(out)
    if res is None:
        pytest.skip("unexpected ncu CSV layout: " + out[start:start + 300].replace("\n", " | "))
    n_rows, agg, col = res
    total, top = ncu_summary.summarize(agg, col, 14)
    warnings.warn("HZPERF ncu_total " + json.dumps({"kernels": n_rows, "distinct": len(agg), "sum_of_durations_us": round(total / 1e3, 1),
                                                   "note": "eager step, serialised under the profiler: shares, not a step time"}))
    for row in top:
        warnings.warn("HZPERF ncu " + json.dumps(row))
    assert n_rows > 50
