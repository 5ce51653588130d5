"""CPU unit tests: contracts (CLI, CSV schema, partitioner, sampler, shard math, schedules) and the
numerical oracle backend against torchvision (SURVEY §4 "Unit" tier)."""
import argparse
import json
import os

import numpy as np
import pandas as pd
import pytest
import torch

from horizonml_b200 import ops
from horizonml_b200.config import TrainConfig, add_train_flags, config_from_args
from horizonml_b200.data import BatchLoader, ShardedSampler, SyntheticCIFAR, build_dataset
from horizonml_b200.metrics import (EXT_COLUMNS, REF_COLUMNS_BW, REF_COLUMNS_DP, EpochRecorder,
                                    merge_worker_csvs, ref_columns)
from horizonml_b200.models.flat import FlatAdam, FlatParams
from horizonml_b200.models.partition import boundary_shape, partition_blocks
from horizonml_b200.models.resnet import resnet18
from horizonml_b200.parallel.pp import one_f_one_b
from horizonml_b200.parallel.tp import padded_classes, shard_range


@pytest.fixture(autouse=True)
def _torch_backend():
    ops.set_backend("torch")
    yield


# ---------------------------------------------------------------------------------- CLI / config
@pytest.mark.parametrize("strategy", ["data", "layer", "tensor"])
def test_cli_reference_flags_and_defaults(strategy):
    p = add_train_flags(argparse.ArgumentParser(), strategy)
    a = p.parse_args([])
    assert (a.world_size, a.epochs, a.sample_size) == (5, 5, 1000)        # reference defaults
    cfg = config_from_args(p.parse_args(["--world_size", "2", "--epochs", "1", "--sample_size", "128"]), strategy)
    assert (cfg.world_size, cfg.epochs, cfg.sample_size, cfg.strategy) == (2, 1, 128, strategy)
    assert cfg.batch_size == 64 and cfg.lr == 1e-3                         # reference literals
    assert cfg.resolved_logs_dir() == {"data": "data_parallel_logs", "layer": "model_parallel_logs",
                                       "tensor": "tensor_parallel_logs"}[strategy]
    assert TrainConfig.from_json(cfg.to_json()) == cfg


def test_main_cli_flags():
    from horizonml_b200 import bench_suite
    assert bench_suite.FIGURES == ["accuracy", "loss", "training_time", "compute_vs_comm", "cpu_utilization",
                                   "memory_usage", "idle_time", "overall_performance"]


# ---------------------------------------------------------------------------------- partitioner
def test_partition_matches_reference_rule():
    assert partition_blocks(1) == [(0, 4)]
    assert partition_blocks(2) == [(0, 2), (3, 4)]
    assert partition_blocks(3) == [(0, 1), (2, 3), (4, 4)]
    assert partition_blocks(4) == [(0, 1), (2, 2), (3, 3), (4, 4)]
    assert partition_blocks(5) == [(i, i) for i in range(5)]
    with pytest.raises(ValueError):
        partition_blocks(6)


def test_boundary_shapes_match_survey_appendix_b():
    # SURVEY App. B (B=64): stem/layer1 [64,64,8,8], layer2 [64,128,4,4], layer3 [64,256,2,2]
    assert boundary_shape(0, 64) == (64, 64, 8, 8)
    assert boundary_shape(1, 64) == (64, 64, 8, 8)
    assert boundary_shape(2, 64) == (64, 128, 4, 4)
    assert boundary_shape(3, 64) == (64, 256, 2, 2)
    m = resnet18(10, seed=0).eval()
    x = torch.randn(4, 3, 32, 32).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for b in range(4):
            x = m.atomic_blocks()[b](x)
            assert tuple(x.shape) == boundary_shape(b, 4)


# ---------------------------------------------------------------------------------- sampler / data
def test_sharded_sampler_is_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    n, ws = 103, 4
    for rank in range(ws):
        ours = ShardedSampler(n, ws, rank, shuffle=False)
        ref = DistributedSampler(list(range(n)), num_replicas=ws, rank=rank, shuffle=False)
        assert list(ours.indices()) == list(iter(ref))
    allidx = np.concatenate([ShardedSampler(n, ws, r, shuffle=True, seed=3).indices() for r in range(ws)])
    assert set(allidx.tolist()) == set(range(n)) and len(allidx) == 104
    s = ShardedSampler(n, ws, 0, shuffle=True, seed=3)
    a = s.indices().copy(); s.set_epoch(1)
    assert not np.array_equal(a, s.indices())


def test_synthetic_dataset_is_seeded_and_cifar_shaped():
    a, la = build_dataset(256, True, "./data", 7)
    b, lb = build_dataset(256, True, "./data", 7)
    assert a.shape == (256, 32, 32, 3) and a.dtype == np.uint8 and np.array_equal(a, b) and np.array_equal(la, lb)
    assert set(np.unique(la).tolist()) <= set(range(10))
    ld = BatchLoader(a, la, 64, "cpu")
    xs = [x for x, _ in ld]
    assert len(ld) == 4 and xs[0].shape == (64, 3, 32, 32) and xs[0].is_contiguous(memory_format=torch.channels_last)
    assert abs(float(xs[0].mean())) < 1.0      # Normalize(0.5, 0.5)


# ---------------------------------------------------------------------------------- CSV schema
def test_csv_schema_matches_reference(tmp_path):
    assert REF_COLUMNS_DP == ["epoch", "loss", "accuracy", "epoch_time", "avg_step_time", "compute_time",
                              "comm_time", "idle_time", "avg_cpu", "avg_memory", "grad_divergence"]
    assert REF_COLUMNS_BW[-2:] == ["avg_bandwidth", "grad_divergence"]
    for strat in ("data", "layer", "tensor"):
        rec = EpochRecorder(strat, 0, str(tmp_path / strat), 128)
        rec.total_compute, rec.total_comm = 1.0, 2.0
        rec.end_epoch(1, 2.3, 10.0, 1.5, [0.1, 0.2], avg_bandwidth=123.0)
        rec.end_epoch(2, 2.0, 20.0, 1.4, [0.1])
        df = pd.read_csv(rec.path)
        assert list(df.columns) == ref_columns(strat) + EXT_COLUMNS
        assert os.path.basename(rec.path) == "worker_0_samples_128.csv"
        assert df["compute_time"].tolist() == [1.0, 1.0]      # cumulative convention (Q8)
    for r in range(2):
        EpochRecorder("data", r, str(tmp_path / "m"), 64).end_epoch(1, 1.0, 1.0, 1.0, [1.0])
    comb = merge_worker_csvs(str(tmp_path / "m"), 2, 64, 9.5)
    assert set(comb["worker"]) == {0, 1} and (comb["total_training_time"] == 9.5).all()
    assert os.path.exists(tmp_path / "m" / "combined_results_64.csv")


# ---------------------------------------------------------------------------------- TP shard math
@pytest.mark.parametrize("ws,kpad", [(1, 10), (2, 10), (4, 12), (5, 10), (8, 16)])
def test_tp_class_padding(ws, kpad):
    assert padded_classes(10, ws) == kpad        # reference truncates to 8 logits at ws=4/8 (Q5)
    spans = [shard_range(kpad, ws, r) for r in range(ws)]
    assert spans[0][0] == 0 and spans[-1][1] == kpad
    assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))


# ---------------------------------------------------------------------------------- 1F1B schedule
@pytest.mark.parametrize("S,M", [(4, 4), (4, 8), (2, 1), (5, 3), (3, 6)])
def test_one_f_one_b_schedule(S, M):
    for s in range(S):
        acts = one_f_one_b(s, S, M)
        assert sorted(a for a in acts if a[0] == "F") == [("F", i) for i in range(M)]
        assert sorted(a for a in acts if a[0] == "B") == [("B", i) for i in range(M)]
        for i in range(M):
            assert acts.index(("F", i)) < acts.index(("B", i))
        inflight = mx = 0
        for a in acts:
            inflight += 1 if a[0] == "F" else -1
            mx = max(mx, inflight)
        assert mx == min(S - s, M)       # 1F1B memory bound: at most (S - stage) live micro-batches


# ---------------------------------------------------------------------------------- model vs torchvision
def test_resnet18_matches_torchvision_forward_backward():
    import torchvision
    ours = resnet18(10, seed=1).train()
    ref = torchvision.models.resnet18(weights=None, num_classes=10).train()
    sd = {k: v.detach().clone().contiguous() for k, v in ours.state_dict().items()}
    ref.load_state_dict(sd)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(16, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (16,), generator=g)
    flat = FlatParams(list(ours.named_parameters()), "cpu", torch.float32)
    loss, correct = ours.forward_loss(x.contiguous(memory_format=torch.channels_last), y)
    loss.backward()
    out = ref(x)
    lref = torch.nn.functional.cross_entropy(out, y)
    lref.backward()
    assert abs(loss.item() - lref.item()) < 1e-4
    assert correct.item() == (out.argmax(1) == y).sum().item()
    refp = dict(ref.named_parameters())
    for n, p in ours.named_parameters():
        g, gr = p.main_grad, refp[n].grad  # noqa: F841
        assert (g - gr).abs().max() <= 2e-3 * gr.abs().max() + 1e-5, (n, (g - gr).abs().max(), gr.abs().max())
    # running statistics follow nn.BatchNorm2d
    assert torch.allclose(ours.bn1.running_var, ref.bn1.running_var, atol=1e-5)
    # eval-mode logits
    ours.eval(); ref.eval()
    with torch.no_grad():
        assert torch.allclose(ours(x.contiguous(memory_format=torch.channels_last)), ref(x), atol=1e-4)


def test_flat_params_layout_and_adam_matches_torch_optim():
    m = resnet18(10, seed=2).train()
    ref = resnet18(10, seed=2).train()
    flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32, bucket_cap_mb=25.0)
    assert flat.names[0] == "fc.bias" and flat.names[-1] == "conv1.weight"        # reverse (backward) order
    assert sum(p.numel() for p in m.parameters()) == 11181642
    assert all(o % 64 == 0 for o in flat.offsets)
    assert flat.buckets[0].start == 0 and flat.buckets[-1].end == flat.total
    assert all(a.end == b.start for a, b in zip(flat.buckets, flat.buckets[1:]))
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert torch.equal(p.detach(), q.detach()), n                              # values survive flattening
    opt, ropt = FlatAdam(flat, lr=1e-3), torch.optim.Adam(ref.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(4, 3, 32, 32, generator=gen).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (4,), generator=gen)
    for it in range(2):
        flat.begin_step()
        m.forward_loss(x, y)[0].backward()
        opt.step()
        ropt.zero_grad()
        ref.forward_loss(x, y)[0].backward()          # plain path: accumulates into .grad
        ropt.step()
        # step 1 is bit-comparable (1 ulp); step 2 sees the batch-4 BatchNorm amplify that ulp
        for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            d = (p.detach() - q.detach()).abs()
            if it == 0:
                assert d.max() <= 1e-6, (it, n)
            else:   # Adam's first steps are sign-like: a flipped tiny gradient moves a weight by ~2*lr
                assert d.mean() <= 2e-5 and d.max() <= 4.1e-3, (it, n, d.mean(), d.max())


def test_microbatch_accumulation_equals_full_batch_grad_for_bn_free_path():
    """main_grad accumulate semantics: two half-batches with loss_scale 0.5 accumulate (BN stats are
    per micro-batch, so compare against the same micro-batched oracle)."""
    m = resnet18(10, seed=3).train()
    flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32)
    x = torch.randn(8, 3, 32, 32).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,))
    flat.begin_step()
    for xs, ys in zip(x.split(4), y.split(4)):
        m.forward_loss(xs.contiguous(memory_format=torch.channels_last), ys, loss_scale=0.5)[0].backward()
    acc = flat.grad.clone()
    parts = []
    for xs, ys in zip(x.split(4), y.split(4)):
        flat.begin_step()
        m.forward_loss(xs.contiguous(memory_format=torch.channels_last), ys, loss_scale=0.5)[0].backward()
        parts.append(flat.grad.clone())
    assert torch.allclose(acc, parts[0] + parts[1], atol=1e-5)


def test_checkpoint_roundtrip(tmp_path):
    from horizonml_b200 import checkpoint
    m = resnet18(10, seed=4).train()
    flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32)
    opt = FlatAdam(flat)
    x = torch.randn(4, 3, 32, 32).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (4,))
    m.forward_loss(x, y)[0].backward(); opt.step()
    checkpoint.save(str(tmp_path), "dp", m, opt, 3, 17)
    m2 = resnet18(10, seed=99).train()
    flat2 = FlatParams(list(m2.named_parameters()), "cpu", torch.float32)
    opt2 = FlatAdam(flat2)
    payload = checkpoint.load(str(tmp_path), "dp", m2, opt2)
    assert payload["epoch"] == 3 and payload["global_step"] == 17
    assert torch.equal(flat2.master, flat.master) and torch.equal(opt2.m, opt.m) and opt2.step_t.item() == 1
    assert torch.equal(m2.bn1.running_mean, m.bn1.running_mean)


def test_bench_suite_summaries_degrade_without_matplotlib(tmp_path):
    from horizonml_b200.bench_suite import generate_comparison_graphs, radar_scores, summarize
    rows = []
    for w in range(2):
        for e in (1, 2):
            rows.append({"epoch": e, "loss": 2.0 / e, "accuracy": 10.0 * e, "epoch_time": 1.0, "avg_step_time": .1,
                         "compute_time": 1.0 * e, "comm_time": 2.0 * e, "idle_time": .1, "avg_cpu": 50, "avg_memory": 100,
                         "grad_divergence": 0, "worker": w, "total_training_time": 5.0, "images_per_sec": 100})
    df = pd.DataFrame(rows)
    mp_df = df.copy(); mp_df.loc[mp_df["worker"] == 0, ["loss", "accuracy"]] = 0      # non-last pipeline ranks write 0
    results = {"data_parallel": {64: df}, "model_parallel": {64: mp_df}, "tensor_parallel": {64: None}}
    s = summarize(results)
    mp_curve = [c for c in s["curves"] if c["strategy"] == "model_parallel" and c["epoch"] == 2][0]
    assert mp_curve["accuracy"] == 20.0                   # last rank only, true world size (Q12)
    assert radar_scores(s["bars"], 64)["data_parallel"]["Accuracy"] == 1.0
    written = generate_comparison_graphs(results, str(tmp_path))
    assert os.path.exists(tmp_path / "benchmark_summary.json")
    for fig in ("accuracy", "loss", "training_time", "compute_vs_comm", "cpu_utilization", "memory_usage",
                "idle_time", "overall_performance"):
        assert os.path.exists(tmp_path / f"{fig}_comparison.csv"), fig


def test_native_extension_builds_and_exports_symbols():
    """The sm_100a extension must be built in-tree (nvcc cross-compiles without a GPU)."""
    from horizonml_b200.ops import _ext
    if _ext._nvcc() is None and not os.path.exists(_ext._SO):
        pytest.skip("no nvcc and no prebuilt extension")
    mod = _ext.load(required=True)
    for sym in ("conv_fwd", "conv_dgrad", "conv_wgrad", "bn_act_fwd", "bn_act_bwd", "head_fwd_bwd", "adam_step",
                "PeerComm", "maxpool_fwd"):
        assert hasattr(mod, sym), sym


def test_dead_tap_masks_are_exact():
    """Conv taps that only ever see padding get an exactly-zero gradient (so optimizer / all-reduce may skip
    them): 62 % of ResNet-18's parameters at 32×32 (layer4 on 1×1 maps)."""
    m = resnet18(10, seed=0).train()
    masks = m.live_tap_masks(32)
    assert set(masks) == {"layer4.0.conv1.weight", "layer4.0.conv2.weight", "layer4.1.conv1.weight",
                          "layer4.1.conv2.weight"}
    assert int(masks["layer4.1.conv2.weight"].sum()) == 512 * 512          # centre tap only
    assert int(masks["layer4.0.conv1.weight"].sum()) == 512 * 256 * 4      # 4 of 9 taps reach the 2×2 input
    flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32, live_masks=masks)
    assert 0.36 < flat.live_fraction < 0.40
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 32, 32, generator=g).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,), generator=g)
    flat.begin_step()
    m.forward_loss(x, y)[0].backward()
    dead = torch.ones(flat.total // 64, dtype=torch.bool)
    dead[flat.live_blocks.long()] = False
    assert flat.grad.view(-1, 64)[dead].abs().max().item() == 0.0
    assert m.live_tap_masks(224) == {}                                       # no dead taps at ImageNet size
    # per-bucket lists are relative to the bucket and cover exactly the live blocks
    tot = sum((b.end - b.start) // 64 if lv is None else lv.numel() for b, lv in zip(flat.buckets, flat.bucket_live))
    assert tot == flat.live_blocks.numel()


def test_gradlink_fusion_equals_plain_autograd(monkeypatch):
    """Residual-gradient hand-off (ops.GradLink): folding the identity/downsample-branch gradient into the
    other branch's dgrad must give the same parameter and input gradients as autograd's own accumulation."""
    from horizonml_b200.models import resnet as R
    gen = torch.Generator().manual_seed(11)
    x0 = torch.randn(8, 3, 32, 32, generator=gen).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,), generator=gen)
    grads = {}
    for fuse in (True, False):
        monkeypatch.setattr(R, "_FUSE_RESADD", fuse)
        m = resnet18(10, seed=4).train()
        flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32)
        flat.begin_step()
        x = x0.clone().requires_grad_(True)
        m.forward_loss(x, y)[0].backward()
        grads[fuse] = (flat.grad.clone(), x.grad.clone())
    for a, b in zip(grads[True], grads[False]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), (a - b).abs().max()


def test_bucketwise_adam_equals_whole_buffer_adam():
    """FlatAdam.step_bucket over all buckets (any order of buckets, first one bumps the step) == FlatAdam.step."""
    outs = []
    for bucketwise in (False, True):
        m = resnet18(10, seed=6).train()
        flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32, bucket_cap_mb=4.0,
                          live_masks=m.live_tap_masks(32))
        assert len(flat.buckets) > 4
        opt = FlatAdam(flat, lr=1e-3)
        prev = torch.zeros_like(flat.grad)
        acc = torch.zeros(())
        gen = torch.Generator().manual_seed(3)
        diffs = []
        for it in range(3):
            flat.grad.copy_(torch.randn(flat.total, generator=gen) * 0.01)
            if bucketwise:
                order = list(range(len(flat.buckets)))
                if it == 1:
                    order = order[::-1]
                for k, b in enumerate(order):
                    opt.step_bucket(b, k == 0, diff_out=acc, prev_grad=prev)
                diffs.append(float(acc))
            else:
                diffs.append(float(opt.step(prev_grad=prev)))
            assert float(flat.grad.abs().max()) == 0.0          # cleared by the optimizer pass
        outs.append((flat.master.clone(), opt.m.clone(), opt.v.clone(), float(opt.step_t), diffs))
    a, b = outs
    assert a[3] == b[3] == 3.0
    for i in range(3):
        assert torch.allclose(a[i], b[i], rtol=1e-6, atol=1e-8)
    assert all(abs(x - y) <= 1e-4 * abs(x) for x, y in zip(a[4], b[4]))


def test_device_mesh_layout():
    """rank = (dp*PP + pp)*TP + tp: TP ranks adjacent, DP outermost; neighbours and group membership are consistent."""
    from horizonml_b200.parallel.mesh import DeviceMesh
    world, dp, pp, tp = 8, 2, 2, 2
    seen = set()
    for r in range(world):
        m = DeviceMesh(world, r, dp=dp, pp=pp, tp=tp, create_groups=False)
        c = m.coord
        assert m.rank_of(c.dp, c.pp, c.tp) == r
        seen.add((c.dp, c.pp, c.tp))
        assert r in m.dp_ranks() and r in m.pp_ranks() and r in m.tp_ranks()
        assert len(m.dp_ranks()) == dp and len(m.pp_ranks()) == pp and len(m.tp_ranks()) == tp
        assert m.tp_ranks() == list(range(r - c.tp, r - c.tp + tp))                   # adjacent GPUs
        prev, nxt = m.pp_neighbours()
        assert (prev is None) == (c.pp == 0) and (nxt is None) == (c.pp == pp - 1)
        if nxt is not None:
            assert DeviceMesh(world, nxt, dp, pp, tp, create_groups=False).pp_neighbours()[0] == r
    assert len(seen) == world
    with pytest.raises(ValueError):
        DeviceMesh(8, 0, dp=3, pp=2, tp=1, create_groups=False)


def test_hybrid_cli_flags():
    import argparse
    from horizonml_b200.config import add_train_flags, config_from_args
    p = argparse.ArgumentParser()
    add_train_flags(p, "layer")
    cfg = config_from_args(p.parse_args(["--world_size", "8", "--dp_replicas", "2", "--overlap_adam",
                                         "--no_region_probe", "--bucket_by_live", "--live_bucket_mb", "4"]), "layer")
    assert cfg.dp_replicas == 2 and cfg.overlap_adam and not cfg.region_probe and cfg.bucket_by_live
    assert cfg.live_bucket_mb == 4.0 and TrainConfig().dp_replicas == 1 and TrainConfig().region_probe


def test_dp_engine_bucketwise_adam_equals_whole_adam():
    """DPEngine with --overlap_adam (Adam per bucket from the reducer's post_bucket hook) == the single fused pass."""
    from horizonml_b200 import ops
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.dp import DPEngine
    ops.set_backend("torch")
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(8, 3, 32, 32, generator=gen).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (8,), generator=gen)
    out = []
    for bw in (False, True):
        cfg = TrainConfig(strategy="data", world_size=1, device="cpu", dtype="fp32", backend="torch", quiet=True,
                          overlap_adam=bw, bucket_mb=4.0, seed=5)
        eng = DPEngine(cfg, Runtime(0, 1, torch.device("cpu"), torch.float32, "torch", "none"))
        assert eng.bucket_adam == bw and (eng.reducer is not None) == bw
        for _ in range(3):
            eng.step(x, y)
        s = eng.stats.read_and_reset()
        out.append((eng.flat.master.clone(), eng.opt.m.clone(), float(eng.opt.step_t), s["loss_sum"], s["grad_div_sum"]))
    a, b = out
    assert a[2] == b[2] == 3.0
    assert torch.allclose(a[0], b[0], rtol=1e-6, atol=1e-8) and torch.allclose(a[1], b[1], rtol=1e-6, atol=1e-9)
    assert abs(a[3] - b[3]) < 1e-5 and abs(a[4] - b[4]) <= 1e-4 * max(abs(a[4]), 1e-12)


def test_svgplot_figures_are_wellformed_xml():
    import xml.etree.ElementTree as ET
    from horizonml_b200 import svgplot
    figs = [
        svgplot.line_chart({"Data Parallel (1000 samples)": [(1, 2.0), (2, 1.0), (3, 0.5)], "b & <c>": [(1, 3.0)]},
                           "Loss Comparison", "Epoch", "Loss"),
        svgplot.line_chart({}, "empty", "x", "y"),
        svgplot.grouped_bars(["1000", "10000"], {"Data Parallel": [1.0, 9.5], "Tensor Parallel": [2.0, 20.0]},
                             "Average Epoch Time (s)", "Sample size", "s"),
        svgplot.grouped_bars(["DP 1000", "TP 1000"], {"Compute": [1.0, 2.0], "Communication": [0.5, 4.0]},
                             "Compute vs Communication", "x", "s", stacked=True),
        svgplot.radar({"Data Parallel": {"Accuracy": 1.0, "Training Speed": 0.3, "Idle Time": 0.0},
                       "Model Parallel": {"Accuracy": 0.5, "Training Speed": 1.2, "Idle Time": -0.1}}, "Overall"),
    ]
    for svg in figs:
        root = ET.fromstring(svg)                       # raises on malformed XML (escaping of & < > included)
        assert root.tag.endswith("svg") and "nan" not in svg


def test_sharded_adam_single_process_matches_flat_adam():
    """ZeRO-1 optimizer with a world of one == FlatAdam (no collectives involved)."""
    from horizonml_b200.parallel.zero import ShardedFlatAdam
    outs = []
    for sharded in (False, True):
        m = resnet18(10, seed=3).train()
        flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32)
        opt = ShardedFlatAdam(flat, lr=1e-3) if sharded else FlatAdam(flat, lr=1e-3)
        gen = torch.Generator().manual_seed(1)
        prev = torch.zeros_like(flat.grad)
        d = None
        for _ in range(2):
            flat.grad.copy_(torch.randn(flat.total, generator=gen) * 0.01)
            d = opt.step(prev_grad=prev)
        outs.append((flat.master.clone(), float(d)))
        assert float(flat.grad.abs().max()) == 0.0
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=1e-8) and abs(outs[0][1] - outs[1][1]) < 1e-6 * outs[0][1]
    sd = opt.state_dict()
    assert sd["m"].numel() == flat.total and float(sd["step"]) == 2.0


def test_real_cifar_batches_reader(tmp_path):
    """--real_data path: CIFAR-10 'python version' pickles (the files torchvision.datasets.CIFAR10 downloads in the
    reference, data_parallel_train.py:49-57) → uint8 NHWC images + int64 labels, seeded subset."""
    import pickle
    from horizonml_b200.data import CIFAR10Files, build_dataset
    base = tmp_path / "cifar-10-batches-py"
    base.mkdir()
    rng = np.random.default_rng(0)
    allx, ally = [], []
    for i in range(1, 6):
        x = rng.integers(0, 256, size=(20, 3072), dtype=np.uint8)
        y = rng.integers(0, 10, size=20).tolist()
        with open(base / f"data_batch_{i}", "wb") as fh:
            pickle.dump({"data": x, "labels": y}, fh)
        allx.append(x); ally += y
    ds = CIFAR10Files(str(tmp_path))
    assert len(ds) == 100 and ds.images.shape == (100, 32, 32, 3) and ds.images.dtype == np.uint8
    x0 = np.concatenate(allx)[7].reshape(3, 32, 32)                       # stored planar CHW → delivered HWC
    assert np.array_equal(ds.images[7], x0.transpose(1, 2, 0)) and ds.labels[7] == ally[7]
    a = build_dataset(30, False, str(tmp_path), seed=5)
    b = build_dataset(30, False, str(tmp_path), seed=5)
    assert a[0].shape == (30, 32, 32, 3) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    with pytest.raises(FileNotFoundError):
        CIFAR10Files(str(tmp_path / "nowhere"))


def test_host_tiling_logic(tmp_path):
    """csrc/tests/host_logic_test.cu: tap elimination, stride-2 parity-view addressing and 128-row tiling of the
    implicit-GEMM convolution, checked against the convolution definition on the host (no GPU, no kernel launch)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_logic_test")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                        os.path.join(root, "csrc"), "-o", exe, os.path.join(root, "csrc", "tests", "host_logic_test.cu")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "host tiling logic ok" in r.stdout, r.stdout + r.stderr
