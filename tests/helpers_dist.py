"""Workers for the multi-process CPU (gloo) tests; importable by spawned children."""
import os
import sys
import traceback

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)


def _batch(n=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, 32, 32, generator=g).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 10, (n,), generator=g)
    return x, y


def run(fn, world, port, out_dir):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, out_dir)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():
            p.terminate()
    errs = [open(os.path.join(out_dir, f)).read() for f in sorted(os.listdir(out_dir)) if f.startswith("err")]
    assert codes == [0] * world, f"exit codes {codes}\n" + "\n".join(errs)


def _entry(fn, rank, world, port, out_dir):
    try:
        _init(rank, world, port)
        if os.environ.get("HZ_BN_BWD_IN_DGRAD") == "1":       # the flag is read at import: make sure the workers have it
            import horizonml_b200.models.resnet as _R
            assert _R._BN_BWD_IN_DGRAD
        globals()[fn](rank, world, out_dir)
        dist.barrier()
        dist.destroy_process_group()
    except BaseException as e:  # noqa: BLE001
        with open(os.path.join(out_dir, f"err{rank}.txt"), "w") as fh:
            fh.write("".join(traceback.format_exception(type(e), e, e.__traceback__)))
        os._exit(1)


# ------------------------------------------------------------------------------------------------
def dp_equivalence(rank, world, out_dir):
    """DP(W) gradient == average over ranks of per-shard gradients (per-rank BN, like DDP)."""
    from horizonml_b200 import ops
    from horizonml_b200.models.flat import FlatParams
    from horizonml_b200.models.resnet import resnet18
    from horizonml_b200.parallel.comm import TorchDistAllReduce
    from horizonml_b200.parallel.dp import GradReducer
    ops.set_backend("torch")
    x, y = _batch(8 * world)
    model = resnet18(10, seed=5).train()
    flat = FlatParams(list(model.named_parameters()), "cpu", torch.float32, bucket_cap_mb=8.0)
    red = GradReducer(flat, TorchDistAllReduce(), overlap=False)
    assert len(flat.buckets) >= 3
    xs, ys = x[rank * 8:(rank + 1) * 8], y[rank * 8:(rank + 1) * 8]
    flat.begin_step(); red.begin_step()
    model.forward_loss(xs.contiguous(memory_format=torch.channels_last), ys)[0].backward()
    red.finish()
    got = flat.grad.clone()
    # oracle: every shard's gradient computed locally, then averaged
    ref = torch.zeros_like(got)
    m2 = resnet18(10, seed=5).train()
    f2 = FlatParams(list(m2.named_parameters()), "cpu", torch.float32, bucket_cap_mb=8.0)
    for r in range(world):
        f2.begin_step()
        m2.forward_loss(x[r * 8:(r + 1) * 8].contiguous(memory_format=torch.channels_last), y[r * 8:(r + 1) * 8])[0].backward()
        ref += f2.grad / world
    assert torch.allclose(got, ref, atol=1e-5), (got - ref).abs().max()
    assert red.bytes_last_step == flat.total * 4


def zero1_equivalence(rank, world, out_dir):
    """ZeRO-1 (reduce-scatter → sharded Adam → all-gather) == replicated Adam after all-reduce, step for step."""
    from horizonml_b200 import ops
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.dp import DPEngine
    ops.set_backend("torch")
    rt = Runtime(rank, world, torch.device("cpu"), torch.float32, "torch", "gloo")
    x, y = _batch(8, seed=20 + rank)                      # a different shard per rank
    res = []
    for zero in (False, True):
        cfg = TrainConfig(strategy="data", world_size=world, device="cpu", dtype="fp32", backend="torch", quiet=True,
                          zero1=zero, seed=9, allreduce="nccl")
        eng = DPEngine(cfg, rt)
        assert eng.zero1 == zero
        if zero:
            assert eng.opt.m.numel() * world == eng.flat.total and eng.ar is None     # moments are 1/W of the model
        for _ in range(3):
            eng.step(x, y)
        s = eng.stats.read_and_reset()
        n = eng.flat.numel()
        full = eng.opt.gather_state() if zero else eng.opt.state_dict()
        res.append((eng.flat.master.clone(), full["m"], s["loss_sum"], s["grad_div_sum"], float(full["step"])))
    a, b = res
    lim = min(a[0].numel(), b[0].numel())                 # the sharded store is padded to a multiple of 64*W
    assert a[4] == b[4] == 3.0
    assert torch.allclose(a[0][:lim], b[0][:lim], rtol=1e-5, atol=1e-7), (a[0][:lim] - b[0][:lim]).abs().max()
    assert torch.allclose(a[1][:lim], b[1][:lim], rtol=1e-5, atol=1e-8)
    assert abs(a[2] - b[2]) < 1e-4 and abs(a[3] - b[3]) <= 1e-3 * max(abs(a[3]), 1e-12)
    # every rank holds the same parameters after the all-gather
    ref = eng.flat.master.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, eng.flat.master)


def zero1_fused_equivalence(rank, world, out_dir):
    """The per-bucket ZeRO-1 optimizer (parallel/zero.py FusedShardedAdam — on CPU its torch.distributed form: same
    wire-order sharding, same bucket loop, same state layout as the one-kernel GPU path) == replicated Adam after
    all-reduce, step for step, with dead-tap compaction on; moments are 1/W of the model; a checkpoint round trip
    through the FlatAdam-compatible gathered state restores the shards."""
    from horizonml_b200 import ops
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.dp import DPEngine
    ops.set_backend("torch")
    rt = Runtime(rank, world, torch.device("cpu"), torch.float32, "torch", "gloo")
    x, y = _batch(8, seed=30 + rank)
    res = []
    for zero in (False, True):
        cfg = TrainConfig(strategy="data", world_size=world, device="cpu", dtype="fp32", backend="torch", quiet=True,
                          zero1=zero, zero1_impl="fused", seed=9, allreduce="nccl", bucket_layout="layers")
        eng = DPEngine(cfg, rt)
        assert eng.zero1 == zero and eng.zero_fused == zero
        if zero:
            assert len(eng.flat.buckets) == 4 and eng.ar is None and not eng.opt.native
            n_wire = sum(i.numel() for i in eng.opt.idx)
            assert n_wire <= eng.flat.total and eng.opt.state_numel <= 2 * (n_wire // world + 8 * len(eng.flat.buckets))
        for _ in range(3):
            eng.step(x, y)
        s = eng.stats.read_and_reset()
        full = eng.opt.gather_state() if zero else eng.opt.state_dict()
        res.append((eng.flat.master.clone(), full["m"], full["v"], s["loss_sum"], s["grad_div_sum"], float(full["step"])))
    a, b = res
    assert a[5] == b[5] == 3.0
    assert torch.allclose(a[0], b[0], rtol=1e-5, atol=1e-7), (a[0] - b[0]).abs().max()
    assert torch.allclose(a[1], b[1], rtol=1e-5, atol=1e-8) and torch.allclose(a[2], b[2], rtol=1e-5, atol=1e-10)
    assert abs(a[3] - b[3]) < 1e-4 and abs(a[4] - b[4]) <= 1e-3 * max(abs(a[4]), 1e-12)
    ref = eng.flat.master.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, eng.flat.master)                      # every rank holds the same parameters
    # state round trip: gathered (FlatAdam layout) -> shards
    m_before = [t.clone() for t in eng.opt.m]
    sd = eng.opt.gather_state()
    for t in eng.opt.m + eng.opt.v:
        t.fill_(123.0)
    eng.opt.load_state_dict(sd)
    for t0, t1, (lo, cnt) in zip(m_before, eng.opt.m, eng.opt.own):
        assert torch.equal(t0[:cnt], t1[:cnt])


def pp_equivalence(rank, world, out_dir):
    """1F1B over `world` stages with M micro-batches == single-process micro-batched gradients."""
    from horizonml_b200 import ops
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.pp import PPEngine
    from horizonml_b200.models.flat import FlatParams
    from horizonml_b200.models.resnet import resnet18
    from horizonml_b200.parallel.pp import one_f_one_b
    ops.set_backend("torch")
    M = 4
    cfg = TrainConfig(strategy="layer", world_size=world, microbatches=M, seed=11, lr=0.0, grad_divergence=False)
    rt = Runtime(rank, world, torch.device("cpu"), torch.float32, "torch", "gloo")
    eng = PPEngine(cfg, rt)
    x, y = _batch(16, seed=3)
    eng.opt.step = lambda **kw: None          # keep the accumulated gradients for inspection
    eng.step(x, y)
    assert eng.runner.trace == one_f_one_b(rank, world, M)
    # oracle on every rank: same micro-batching, whole model
    m = resnet18(10, seed=11).train()
    flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32)
    flat.begin_step()
    tot = 0.0
    for xs, ys in zip(x.split(4), y.split(4)):
        l, _ = m.forward_loss(xs.contiguous(memory_format=torch.channels_last), ys, loss_scale=0.25)
        l.backward(); tot += l.item()
    ref = {n: p.main_grad for n, p in m.named_parameters()}
    for n, p in zip(eng.flat.names, eng.flat.params):
        assert torch.allclose(p.main_grad, ref[n], atol=2e-5), (n, (p.main_grad - ref[n]).abs().max())
    if eng.is_last:
        s = eng.stats.read_and_reset()
        assert abs(s["loss_sum"] - tot) < 1e-4


def hybrid_dp_pp_equivalence(rank, world, out_dir):
    """DP(2) x PP(world/2) process mesh: every stage's gradients == mean over replicas of the single-process
    micro-batched gradient on that replica's batch."""
    from horizonml_b200 import ops
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.pp import PPEngine
    from horizonml_b200.models.flat import FlatParams
    from horizonml_b200.models.resnet import resnet18
    from horizonml_b200.parallel.mesh import DeviceMesh
    from horizonml_b200.parallel.pp import one_f_one_b
    ops.set_backend("torch")
    M, dp = 4, 2
    mesh = DeviceMesh(world, rank, dp=dp, pp=world // dp)
    assert mesh.rank_of(mesh.coord.dp, mesh.coord.pp, 0) == rank
    assert sorted(mesh.dp_ranks() + mesh.pp_ranks()).count(rank) == 2
    cfg = TrainConfig(strategy="layer", world_size=world, microbatches=M, seed=11, lr=0.0, grad_divergence=False,
                      dp_replicas=dp)
    rt = Runtime(rank, world, torch.device("cpu"), torch.float32, "torch", "gloo")
    eng = PPEngine(cfg, rt, mesh)
    batches = [_batch(16, seed=3 + d) for d in range(dp)]
    x, y = batches[mesh.coord.dp]
    eng.opt.step = lambda **kw: None
    eng.step(x, y)
    assert eng.runner.trace == one_f_one_b(mesh.coord.pp, mesh.pp, M)
    ref = None
    for xb, yb in batches:
        m = resnet18(10, seed=11).train()
        flat = FlatParams(list(m.named_parameters()), "cpu", torch.float32)
        flat.begin_step()
        for xs, ys in zip(xb.split(4), yb.split(4)):
            m.forward_loss(xs.contiguous(memory_format=torch.channels_last), ys, loss_scale=0.25)[0].backward()
        g = {n: p.main_grad.clone() / dp for n, p in m.named_parameters()}
        ref = g if ref is None else {n: ref[n] + g[n] for n in g}
    for n, p in zip(eng.flat.names, eng.flat.params):
        assert torch.allclose(p.main_grad, ref[n], atol=2e-5), (n, (p.main_grad - ref[n]).abs().max())


def hybrid_dp_tp_equivalence(rank, world, out_dir):
    """DP(2) x TP(world/2) mesh through TPEngine: replicated + sharded gradients == mean over replicas of the
    dense model's gradient on that replica's batch."""
    from horizonml_b200 import ops
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.tp import TPEngine
    from horizonml_b200.models.flat import FlatParams
    from horizonml_b200.models.resnet import resnet18
    from horizonml_b200.parallel.mesh import DeviceMesh
    from horizonml_b200.parallel.tp import shard_range, padded_classes
    ops.set_backend("torch")
    dp = 2
    mesh = DeviceMesh(world, rank, dp=dp, tp=world // dp)
    cfg = TrainConfig(strategy="tensor", world_size=world, seed=21, lr=0.0, grad_divergence=False, dp_replicas=dp,
                      cuda_graph=False)
    rt = Runtime(rank, world, torch.device("cpu"), torch.float32, "torch", "gloo")
    eng = TPEngine(cfg, rt, mesh)
    assert eng.comm.world == mesh.tp and eng.comm.rank == mesh.coord.tp
    batches = [_batch(8, seed=4 + d) for d in range(dp)]
    eng.opt_rep.step = lambda **kw: None
    eng.opt_shd.step = lambda **kw: None
    x, y = batches[mesh.coord.dp]
    eng.step(x, y)
    ref = None
    for xb, yb in batches:
        dense = resnet18(10, seed=21).train()
        FlatParams(list(dense.named_parameters()), "cpu", torch.float32)
        dense.forward_loss(xb, yb)[0].backward()
        g = {n: p.main_grad.clone() / dp for n, p in dense.named_parameters()}
        ref = g if ref is None else {n: ref[n] + g[n] for n in g}
    got = dict(eng.model.named_parameters())
    tpw, tpr = mesh.tp, mesh.coord.tp
    kpad = padded_classes(10, tpw)
    lo, hi = shard_range(kpad, tpw, tpr)
    wref = torch.zeros(kpad, 512); wref[:10] = ref["fc.weight"]
    assert torch.allclose(got["fc_weight"].main_grad, wref[lo:hi], atol=2e-5)
    n_rep = 0
    for n, p in got.items():
        dn = n[len("backbone."):] if n.startswith("backbone.") else n
        if not getattr(p, "tp_sharded", False) and dn in ref:
            assert torch.allclose(p.main_grad, ref[dn], atol=2e-5), (n, (p.main_grad - ref[dn]).abs().max())
            n_rep += 1
    assert n_rep > 20
    # a sharded conv (layer4 conv1 is column-parallel: this rank's Cout slice of the dense gradient)
    clo, chi = shard_range(512, tpw, tpr)
    w = got["backbone.layer4.1.conv1.weight"]
    if getattr(w, "tp_sharded", False):
        assert torch.allclose(w.main_grad, ref["layer4.1.conv1.weight"][clo:chi], atol=2e-5)


def tp_equivalence(rank, world, out_dir):
    """TP(W) (column-parallel classifier + channel-parallel layer3/4) == dense model."""
    from horizonml_b200 import ops
    from horizonml_b200.models.flat import FlatParams
    from horizonml_b200.models.resnet import resnet18
    from horizonml_b200.parallel.tp import TensorParallelResNet, TPComm, shard_range, padded_classes
    ops.set_backend("torch")
    x, y = _batch(8, seed=4)
    dense = resnet18(10, seed=21).train()
    fd = FlatParams(list(dense.named_parameters()), "cpu", torch.float32)
    ld, cd = dense.forward_loss(x, y)
    ld.backward()
    ref = {n: p.main_grad.clone() for n, p in dense.named_parameters()}
    comm = TPComm()
    tp = TensorParallelResNet(resnet18(10, seed=21), comm, conv_split=True).train()
    rep, shd = tp.split_params()
    fr = FlatParams(rep, "cpu", torch.float32); fs = FlatParams(shd, "cpu", torch.float32)
    lt, ct = tp.forward_loss(x, y)
    lt.backward()
    assert abs(lt.item() - ld.item()) < 1e-4 and ct.item() == cd.item()
    kpad = padded_classes(10, world)
    lo, hi = shard_range(kpad, world, rank)
    got = dict(tp.named_parameters())
    wref = torch.zeros(kpad, 512); wref[:10] = ref["fc.weight"]
    assert torch.allclose(got["fc_weight"].main_grad, wref[lo:hi], atol=2e-5)
    for lname, cout in (("layer3", 256), ("layer4", 512)):
        clo, chi = shard_range(cout, world, rank)
        for b in (0, 1):
            g1 = got[f"backbone.{lname}.{b}.conv1.weight"].main_grad
            g2 = got[f"backbone.{lname}.{b}.conv2.weight"].main_grad
            assert torch.allclose(g1, ref[f"{lname}.{b}.conv1.weight"][clo:chi], atol=5e-5), lname
            assert torch.allclose(g2, ref[f"{lname}.{b}.conv2.weight"][:, clo:chi], atol=5e-5), lname
            assert torch.allclose(got[f"backbone.{lname}.{b}.bn1.weight"].main_grad,
                                  ref[f"{lname}.{b}.bn1.weight"][clo:chi], atol=5e-5)
    for n in ("conv1.weight", "layer1.0.conv1.weight", "layer2.0.downsample.0.weight", "layer4.1.bn2.weight"):
        assert torch.allclose(got["backbone." + n].main_grad, ref[n], atol=5e-5), n
    assert comm.take_bytes() > 0


def tp_linear_equivalence(rank, world, out_dir):
    """ColumnParallelLinear → RowParallelLinear (the Megatron MLP pattern) == dense Linear → Linear, incl. grads;
    out_features=10 is not divisible by the group size (padding path)."""
    from horizonml_b200.parallel.tp import ColumnParallelLinear, RowParallelLinear, TPComm, shard_range, padded_classes
    comm = TPComm()
    g = torch.Generator().manual_seed(0)
    w1, b1 = torch.randn(64, 32, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1
    w2, b2 = torch.randn(10, 64, generator=g) * 0.1, torch.randn(10, generator=g) * 0.1
    x = torch.randn(8, 32, generator=g, requires_grad=True)
    y = torch.relu(x @ w1.t() + b1) @ w2.t() + b2
    y.square().sum().backward()
    gx_ref = x.grad.clone()
    col = ColumnParallelLinear(32, 64, comm, full_weight=w1, full_bias=b1, gather_output=False)
    row = RowParallelLinear(64, 10, comm, full_weight=w2, full_bias=b2)
    x2 = x.detach().clone().requires_grad_(True)
    y2 = row(torch.relu(col(x2)))
    assert torch.allclose(y2, y.detach(), atol=1e-5)
    y2.square().sum().backward()
    assert torch.allclose(x2.grad, gx_ref, atol=1e-5)
    # gathered column-parallel head with padding: 10 classes over `world` ranks
    head = ColumnParallelLinear(64, 10, comm, full_weight=w2, full_bias=b2, gather_output=True)
    h = torch.randn(8, 64, generator=g)
    assert torch.allclose(head(h), h @ w2.t() + b2, atol=1e-5)
    assert head.weight.shape[0] == padded_classes(10, world) // world
