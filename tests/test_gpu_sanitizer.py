"""compute-sanitizer over the kernel numerics tests: the GPU test run itself executes ``memcheck`` on the
hardware-verified SIMT kernels (BatchNorm forward / backward, max-pool, classifier head, fused Adam) and on one tcgen05 /
TMA convolution, and requires ``ERROR SUMMARY: 0 errors`` — so every GPU test log carries a sanitizer verdict
(`tools/sanitize.sh` runs the full memcheck / synccheck / racecheck passes by hand).

`late`: written after the round's GPU budget was spent.  The sanitizer slows kernels 10-50x and python start-up under it
takes a while, so the default selection is tiny and a run that does not finish inside its time box is a skip, not a
failure; the convolution pass is opt-in (HZ_TEST_SANITIZER=1)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _memcheck(selection: str, files, budget_s: int, exe: str = ""):
    exe = exe or shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    # plain stream order (the sanitizer serialises kernels anyway); lazy module loading: with the eager loading the
    # virtual-rank tests ask for (tests/conftest.py) the tool would instrument every kernel image of the process
    env = dict(os.environ, HZ_PDL="0", CUDA_MODULE_LOADING="LAZY")
    cmd = [exe, "--tool", "memcheck", "--error-exitcode", "9", "--launch-timeout", "60", sys.executable, "-m", "pytest",
           *files, "-q", "-x", "-m", "gpu and not late", "-k", selection, "-p", "no:cacheprovider"]
    import signal
    proc = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True)           # own process group: tool + target die together
    try:
        out, _ = proc.communicate(timeout=budget_s)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except OSError:
            pass
        proc.communicate()
        pytest.skip(f"compute-sanitizer run did not finish within {budget_s} s")
    rc = proc.returncode
    if "ERROR SUMMARY" not in out:
        pytest.skip("compute-sanitizer produced no summary (tool could not attach?): " + out[-400:].replace("\n", " | "))
    import json
    import re
    import warnings
    summ = re.findall(r"ERROR SUMMARY: (\d+) error", out)
    ran = re.findall(r"(\d+) passed", out)
    warnings.warn("HZPERF sanitizer " + json.dumps({
        "tool": "compute-sanitizer memcheck", "selection": selection, "error_summaries": [int(x) for x in summ],
        "tests_passed_under_the_tool": int(ran[-1]) if ran else 0, "exit_code": rc}), UserWarning)
    assert "ERROR SUMMARY: 0 errors" in out and rc == 0, out[-3000:]
    assert " passed" in out, out[-1500:]


@pytest.mark.gpu
@pytest.mark.late(order=10, limit_s=120)     # right after the numerics tests and the ncu report: hardware-verified kernels only
def test_memcheck_clean_on_elementwise_kernels():
    """BatchNorm forward + backward (reduce, apply), max-pool forward / backward, the classifier head: 75 s box."""
    _memcheck("test_bn_act and 64-16 or test_maxpool or test_head and 1", ["tests/test_gpu_kernels.py"], 75)


@pytest.mark.gpu
@pytest.mark.late(order=14, limit_s=700)
@pytest.mark.skipif(os.environ.get("HZ_TEST_SANITIZER", "0") != "1", reason="set HZ_TEST_SANITIZER=1 (several minutes)")
def test_memcheck_clean_on_a_tcgen05_convolution():
    _memcheck("(test_conv_fwd_tcgen05 or test_conv_dgrad or test_conv_wgrad_tcgen05) and cfg0", ["tests/test_gpu_kernels.py"], 600)
