"""Multi-process CPU tests on the fake cluster the reference itself uses — gloo over loopback
(SURVEY §4): every entrypoint at world_size=2 (BASELINE.json config #1), strategy equivalence against
single-process math, clean shutdown, failure propagation."""
import os
import time

import pandas as pd
import pytest

from horizonml_b200.config import TrainConfig
from horizonml_b200.launch import find_free_port
from horizonml_b200.metrics import EXT_COLUMNS, ref_columns
from horizonml_b200.trainers import run_data_parallel, run_model_parallel, run_tensor_parallel

from helpers_dist import run

pytestmark = pytest.mark.slow
FAST = TrainConfig(batch_size=16, device="cpu", quiet=True, watchdog_s=240)


def test_dp_equivalence_gloo(tmp_path):
    run("dp_equivalence", 2, find_free_port(), str(tmp_path))


def test_pp_1f1b_equivalence_gloo(tmp_path):
    run("pp_equivalence", 3, find_free_port(), str(tmp_path))


def test_tp_equivalence_gloo(tmp_path):
    run("tp_equivalence", 2, find_free_port(), str(tmp_path))


@pytest.mark.parametrize("runner,strategy,ws", [(run_data_parallel, "data", 2), (run_model_parallel, "layer", 2),
                                                (run_tensor_parallel, "tensor", 2)])
def test_entrypoints_world2_cpu(tmp_path, runner, strategy, ws):
    t0 = time.time()
    df = runner(ws, 2, 64, logs_dir=str(tmp_path), cfg=FAST)
    took = time.time() - t0
    assert df is not None, "job failed"
    assert took < 200, "shutdown must be clean (the reference hangs until its watchdog fires, Q7)"
    assert list(df.columns) == ref_columns(strategy) + EXT_COLUMNS + ["worker", "total_training_time"]
    assert sorted(df["worker"].unique()) == list(range(ws)) and df["epoch"].max() == 2
    assert abs(df["total_training_time"].iloc[0] - took) < 5          # true wall time, not a watchdog
    assert os.path.exists(tmp_path / "combined_results_64.csv")
    last = df[df["worker"] == ws - 1]
    assert last["loss"].iloc[-1] < last["loss"].iloc[0]                # it trains
    if strategy == "layer":                                             # non-last stages write 0 (reference)
        assert (df[df["worker"] == 0]["loss"] == 0).all()
        assert (df["avg_bandwidth"] > 0).any()


def test_fault_injection_tears_job_down(tmp_path):
    cfg = FAST.replace(inject_fault="1:1")
    t0 = time.time()
    df = run_data_parallel(2, 3, 128, logs_dir=str(tmp_path), cfg=cfg)
    assert df is None
    assert time.time() - t0 < 120, "a dead rank must not leave the job hanging"
    assert os.path.exists(tmp_path / "error_rank1.txt")
    assert "injected fault" in open(tmp_path / "error_rank1.txt").read()
