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


def test_zero1_sharded_optimizer_equivalence_gloo(tmp_path):
    run("zero1_equivalence", 2, find_free_port(), str(tmp_path))


def test_zero1_per_bucket_fused_form_equivalence_gloo(tmp_path):
    run("zero1_fused_equivalence", 2, find_free_port(), str(tmp_path))


def test_pp_1f1b_equivalence_gloo(tmp_path):
    run("pp_equivalence", 3, find_free_port(), str(tmp_path))


def test_equivalences_hold_with_bn_backward_sums_hand_off(tmp_path, monkeypatch):
    """HZ_BN_BWD_IN_DGRAD=1 (BatchNorm-backward sums handed over by the consuming dgrad / max-pool backward,
    ops.BNBackLink) must not change any strategy's gradients: pipeline micro-batches (links per micro-batch, stage
    boundaries without a producer) and tensor-parallel blocks (bn1 <- the row-parallel conv2's local dgrad) against the
    same dense references as above, on the PyTorch-op oracle of the fused kernels."""
    monkeypatch.setenv("HZ_BN_BWD_IN_DGRAD", "1")          # read at import by the spawned workers
    run("pp_equivalence", 3, find_free_port(), str(tmp_path))
    run("tp_equivalence", 2, find_free_port(), str(tmp_path))


def test_hybrid_dp_pp_mesh_equivalence_gloo(tmp_path):
    run("hybrid_dp_pp_equivalence", 4, find_free_port(), str(tmp_path))


def test_hybrid_dp_tp_mesh_equivalence_gloo(tmp_path):
    run("hybrid_dp_tp_equivalence", 4, find_free_port(), str(tmp_path))


def test_tp_equivalence_gloo(tmp_path):
    run("tp_equivalence", 2, find_free_port(), str(tmp_path))


def test_tp_linear_layers_gloo(tmp_path):
    run("tp_linear_equivalence", 4, find_free_port(), str(tmp_path))


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
    if strategy == "data":                                              # region probe fills the extended columns
        assert (df["fwd_ms"] > 0).all() and (df["bwd_ms"] > 0).all() and (df["optimizer_ms"] > 0).all()
        assert (df["allreduce_ms"] > 0).all()
    if strategy == "layer":                                             # non-last stages write 0 (reference)
        assert (df[df["worker"] == 0]["loss"] == 0).all()
        assert (df["avg_bandwidth"] > 0).any()
        assert (df["p2p_ms"] > 0).all()                                 # exposed exchange time per step, every stage


@pytest.mark.parametrize("inner", ["layer", "tensor"])
def test_hybrid_entrypoint_world4_cpu(tmp_path, inner):
    """hybrid_parallel_train.py: 2 replicas x (2-stage pipeline | TP-2 group) on a 4-process gloo mesh."""
    import hybrid_parallel_train as hp
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        rc = hp.main(["--world_size", "4", "--dp_replicas", "2", "--inner", inner, "--epochs", "2",
                      "--sample_size", "64", "--device", "cpu", "--batch_size", "16", "--quiet",
                      "--logs_dir", str(tmp_path / "logs")])
    finally:
        os.chdir(cwd)
    assert rc == 0
    df = pd.read_csv(tmp_path / "logs" / "combined_results_64.csv")
    assert sorted(df["worker"].unique()) == [0, 1, 2, 3] and df["epoch"].max() == 2
    last = df[df["worker"] == 3]
    assert last["loss"].iloc[-1] < last["loss"].iloc[0]


def test_fault_injection_tears_job_down(tmp_path):
    cfg = FAST.replace(inject_fault="1:1")
    t0 = time.time()
    df = run_data_parallel(2, 3, 128, logs_dir=str(tmp_path), cfg=cfg)
    assert df is None
    assert time.time() - t0 < 120, "a dead rank must not leave the job hanging"
    assert os.path.exists(tmp_path / "error_rank1.txt")
    assert "injected fault" in open(tmp_path / "error_rank1.txt").read()


def test_checkpoint_resume_through_trainer(tmp_path):
    """--save_dir / --resume: epoch counter, weights and Adam state survive (absent in the reference)."""
    import torch
    cfg = FAST.replace(save_dir=str(tmp_path / "ck"))
    df1 = run_data_parallel(1, 1, 64, logs_dir=str(tmp_path / "a"), cfg=cfg)
    assert df1 is not None and os.path.exists(tmp_path / "ck" / "ckpt_dp.pt")
    ck = torch.load(tmp_path / "ck" / "ckpt_dp.pt", map_location="cpu", weights_only=False)
    assert ck["epoch"] == 1 and ck["global_step"] == 4 and float(ck["optim"]["step"][0]) == 4
    df2 = run_data_parallel(1, 2, 64, logs_dir=str(tmp_path / "b"), cfg=cfg.replace(resume=str(tmp_path / "ck")))
    assert df2 is not None and df2["epoch"].tolist() == [2]          # resumed: only epoch 2 is run
    assert df2["loss"].iloc[0] < df1["loss"].iloc[0]


def test_legacy_train_entrypoint(tmp_path):
    """train.py: env-var rendezvous, MODEL_TYPE, legacy CSV columns (reference train.py:15-126)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(find_free_port()),
               MODEL_TYPE="resnet", EPOCHS="1", SAMPLE_SIZE="64", BATCH_SIZE="16", DEVICE="cpu", LOGS_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(root, "train.py")], cwd=tmp_path, env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    df = pd.read_csv(tmp_path / "training_logs_worker_0.csv")
    assert list(df.columns) == ["Worker", "Epoch", "Loss", "Accuracy", "Time"] and len(df) == 1


def test_main_benchmark_sweep_cpu(tmp_path):
    """main.py programme: the three strategies in turn + the eight comparison artefacts (reference main.py:17-61,64-390)."""
    from horizonml_b200.bench_suite import generate_comparison_graphs, run_benchmarks
    cfg = FAST.replace(logs_dir=None)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        res = run_benchmarks([64], 2, 1, cfg)
        written = generate_comparison_graphs(res, str(tmp_path / "out"))
    finally:
        os.chdir(cwd)
    assert all(res[s][64] is not None for s in ("data_parallel", "model_parallel", "tensor_parallel"))
    summ = __import__("json").load(open(tmp_path / "out" / "benchmark_summary.json"))
    assert {b["strategy"] for b in summ["bars"]} == {"data_parallel", "model_parallel", "tensor_parallel"}
    assert os.path.exists(tmp_path / "out" / "overall_performance_comparison.csv")
    for fig in ("accuracy", "loss", "training_time", "compute_vs_comm", "cpu_utilization", "memory_usage", "idle_time",
                "overall_performance"):        # the reference's eight figures, drawn without matplotlib
        svg = open(tmp_path / "out" / f"{fig}_comparison.svg").read()
        assert svg.startswith("<svg") and svg.rstrip().endswith("</svg>")
        assert ("Data Parallel" in svg) or fig == "compute_vs_comm"
    # analyze_results.py: the same reports rebuilt offline from the logs the sweep left behind
    from horizonml_b200.bench_suite import analyze_main, load_results
    again = load_results(None, str(tmp_path))
    assert all(again[s][64] is not None and len(again[s][64]) == len(res[s][64]) for s in again)
    assert analyze_main(["--logs_root", str(tmp_path), "--output_dir", str(tmp_path / "out2")]) == 0
    s2 = __import__("json").load(open(tmp_path / "out2" / "benchmark_summary.json"))
    assert {b["strategy"] for b in s2["bars"]} == {"data_parallel", "model_parallel", "tensor_parallel"}
