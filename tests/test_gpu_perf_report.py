"""Round-end measurements taken inside the GPU test run and published through pytest's warnings summary (lines starting
with ``HZPERF``): the only way code written after the round's GPU budget was spent gets a number at all.  The
measurements run in tools/perf_probe.py, one process per section (a kernel that traps cannot take the other sections'
numbers — or this session's CUDA context — with it); device-timed with CUDA events after warm-up, synchronised on both
sides, a 256 MiB L2-flush write between timed launches (the recipe of tools/conv_roofline.py / bench.py).  Nothing here
can fail the run: every section reports what it could measure, or why not.

* steps   — ResNet-18 and MobileNetV2 training step through the DP engine on one GPU (ms/step, images/s, launches/step)
* handoff — the same steps with the BatchNorm-backward sums taken in the dgrad / pool-backward kernels (HZ_BN_BWD_IN_DGRAD)
* bench   — bench.py --gpus 1 (its own timing rules and e2e arm) on the default path and with the hand-offs
* conv    — batch-4096 convolutions: one-tile-per-CTA kernel vs cuDNN vs the persistent kernels (TFLOP/s, fraction of peak)
* bigbatch — ResNet-18 training step at batch 2048: default kernels, PyTorch ops, persistent kernels chosen by the wave rule"""
import os
import signal
import subprocess
import sys
import warnings

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=11)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("section,budget_s", [("steps", 70), ("handoff", 60), ("bench", 90), ("conv", 60), ("bigbatch", 70)])
def test_round_end_perf_report(section, budget_s):
    proc = subprocess.Popen([sys.executable, os.path.join(ROOT, "tools", "perf_probe.py"), section], cwd=ROOT,
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=budget_s)
        note = "" if proc.returncode == 0 else f"exit code {proc.returncode}: " + (err or "")[-300:].replace("\n", " | ")
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except OSError:
            pass
        out, err = proc.communicate()
        note = f"section did not finish within {budget_s} s"
    lines = [ln for ln in (out or "").splitlines() if ln.startswith("HZPERF ")]
    for ln in lines:
        warnings.warn(ln, UserWarning)
    if note:
        warnings.warn(f"HZPERF {section}_incomplete " + note, UserWarning)
    assert lines or note          # (informational: the section either reported numbers or says why not)
