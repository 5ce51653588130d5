"""Tensor-parallel kernels on ONE GPU: `world` *virtual ranks* (parallel/symm.py ``SymmHeap.virtual``) run the real
multi-rank protocol of csrc/tp_fused.cu — bf16 partial slots, arrival counters, rank-ordered peer pulls, parity
double-buffering, epoch counters — concurrently on `world` streams, and every result is compared against a plain
PyTorch fp32 reference of the same op summed over the shards (SURVEY §4 "Collective correctness" + "Strategy
equivalence" tiers, made runnable on the driver's single-GPU box).  The NVSwitch multicast variant of the same
kernels (``multimem.red`` / ``multimem.ld_reduce``) needs real peers: tools/tp_fused_check.py under torchrun.

Also: the dense tcgen05 conv kernels on channel counts that are not multiples of 64 — the 32-channel shards of
layer3 at tensor-parallel size 8 (reference: ``out_features // world_size``, tensor_parallel_train.py:32)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


@pytest.fixture(scope="module")
def nb():
    from horizonml_b200.ops import native_backend
    native_backend.C.install_crash_backtrace()      # a native crash prints its C++ frames instead of dying silently
    return native_backend


@pytest.fixture(scope="module")
def tb():
    from horizonml_b200.ops import torch_backend
    return torch_backend


def _virtual(world, mb=48):
    from horizonml_b200.parallel.symm import SymmHeap
    from horizonml_b200.parallel.tp import FusedTP
    heaps = SymmHeap.virtual(world, DEV, mb << 20)
    return [FusedTP(DEV, heap=h) for h in heaps]


def _run_ranks(fns):
    """fns[r]() on its own stream; returns the results after a device sync (the kernels wait for each other)."""
    streams = [torch.cuda.Stream() for _ in fns]
    torch.cuda.synchronize()
    out = [None] * len(fns)
    for r, fn in enumerate(fns):
        with torch.cuda.stream(streams[r]):
            out[r] = fn()
    torch.cuda.synchronize()
    return out


# ------------------------------------------------------------------------------------------------------------
NARROW = [  # N, Cin, H, W, Cout, R, stride, pad: the layer3 / layer4 shard shapes at tensor-parallel size 8 and 4
    (64, 128, 4, 4, 32, 3, 2, 1), (64, 256, 2, 2, 32, 3, 1, 1), (64, 32, 2, 2, 256, 3, 1, 1),
    (64, 64, 2, 2, 256, 3, 1, 1), (64, 256, 2, 2, 64, 3, 2, 1), (64, 512, 1, 1, 64, 3, 1, 1),
    (64, 64, 1, 1, 512, 3, 1, 1), (64, 128, 4, 4, 96, 3, 1, 1), (64, 32, 4, 4, 32, 1, 1, 0),
]


@pytest.mark.parametrize("cfg", NARROW)
def test_conv_narrow_channels(nb, tb, cfg):
    """fwd (+BN sums) / dgrad (+addend) / wgrad with Cin or Cout below / not a multiple of 64: TMA zero-fills the
    64-wide boxes, the epilogues mask the surplus columns — no cuDNN fallback for tensor-parallel shards."""
    N, Cin, H, W, Cout, R, s, p = cfg
    g = torch.Generator().manual_seed(3)
    x = cl((torch.randn(N, Cin, H, W, generator=g) * 0.5).to(DEV).bfloat16())
    w = cl((torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(DEV).bfloat16())
    Ho, Wo = (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    dy = cl((torch.randn(N, Cout, Ho, Wo, generator=g) * 0.5).to(DEV).bfloat16())
    add = cl((torch.randn(N, Cin, H, W, generator=g) * 0.5).to(DEV).bfloat16())
    fb = dict(nb.FALLBACKS)
    y, stats = nb.conv_fwd(x, w, s, p, True)
    dx = nb.conv_dgrad(dy, w, x.shape, s, p, add)
    dw = torch.zeros(Cout, R, R, Cin, device=DEV).permute(0, 3, 1, 2)           # storage [Cout,R,S,Cin]; dead taps are never written
    nb.conv_wgrad(dy, x, w.shape, s, p, dw, False)
    torch.cuda.synchronize()
    assert dict(nb.FALLBACKS) == fb, "narrow-channel conv fell back to the PyTorch oracle"
    yr, sr = tb.conv_fwd(x.float(), w.float(), s, p, True)
    assert rel_err(y, yr) < 2e-2 and rel_err(stats, sr) < 2e-2
    dxr = tb.conv_dgrad(dy.float(), w.float(), x.shape, s, p) + add.float()
    assert rel_err(dx, dxr) < 2e-2
    dwr = torch.zeros(Cout, R, R, Cin, device=DEV).permute(0, 3, 1, 2)
    tb.conv_wgrad(dy.float(), x.float(), w.shape, s, p, dwr, False)
    assert rel_err(dw, dwr) < 2e-2


# ------------------------------------------------------------------------------------------------------------
# (kind, x_shape [N,Cin,H,W] of the DENSE layer, Cout of the dense layer, stride): kind 0 = row-parallel conv2 forward
# (Cin split), kind 1 = column-parallel conv1 dgrad (Cout split) — layer3 / layer4 of ResNet-18 at batch 64
TP_CASES = [
    (0, (64, 256, 2, 2), 256, 1), (0, (64, 512, 1, 1), 512, 1),
    (1, (64, 256, 2, 2), 256, 1), (1, (64, 128, 4, 4), 256, 2), (1, (64, 256, 2, 2), 512, 2), (1, (64, 512, 1, 1), 512, 1),
]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", TP_CASES)
def test_fused_gemm_allreduce_virtual_ranks(nb, tb, world, case):
    kind, xs, cout, stride = case
    n, cin, h, w = xs
    ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
    fz = _virtual(world)
    g = torch.Generator().manual_seed(11)
    W_full = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    if kind == 0:      # split Cin: a_r = x[:, shard], w_r = W[:, shard]
        x = torch.randn(n, cin, h, w, generator=g) * 0.5
        k = cin // world
        a = [cl(x[:, r * k:(r + 1) * k].to(DEV).bfloat16()) for r in range(world)]
        ws = [cl(W_full[:, r * k:(r + 1) * k].to(DEV).bfloat16()) for r in range(world)]
        ref = sum(tb.conv_fwd(a[r].float(), ws[r].float(), 1, 1, False)[0] for r in range(world))
        shard_x = (n, k, h, w)
        ops_ = [f.allreduce_conv(0, shard_x, (cout, k, 3, 3), 1, 1) for f in fz]
        add = None
    else:              # split Cout: a_r = dy[:, shard], w_r = W[shard]
        dy = torch.randn(n, cout, ho, wo, generator=g) * 0.5
        k = cout // world
        a = [cl(dy[:, r * k:(r + 1) * k].to(DEV).bfloat16()) for r in range(world)]
        ws = [cl(W_full[r * k:(r + 1) * k].to(DEV).bfloat16()) for r in range(world)]
        ref = sum(tb.conv_dgrad(a[r].float(), ws[r].float(), xs, stride, 1) for r in range(world))
        add = cl((torch.randn(*xs, generator=g) * 0.5).to(DEV).bfloat16())
        ref = ref + add.float()
        ops_ = [f.allreduce_conv(1, xs, (k, cin, 3, 3), stride, 1) for f in fz]
    nout = ref.shape[1]
    for it in range(4):                                   # re-launch: epochs, parity slots, counters
        stats = [torch.zeros(2, nout, device=DEV) for _ in range(world)] if kind == 0 else [None] * world
        ys = _run_ranks([(lambda r=r: ops_[r](a[r], ws[r], add, stats[r])) for r in range(world)])
        for r in range(world):
            assert rel_err(ys[r], ref) < 2.5e-2, (r, it)
            assert torch.equal(ys[r], ys[0]), "ranks disagree bit-wise"
            if kind == 0:
                yf = ys[r].float()
                sr = torch.stack([yf.sum(dim=(0, 2, 3)), (yf * yf).sum(dim=(0, 2, 3))])
                assert rel_err(stats[r], sr) < 1e-3


@pytest.mark.parametrize("world", [2, 8])
def test_fused_gemm_reduce_scatter_virtual_ranks(nb, tb, world):
    """mode 2: tile t is reduced and kept by rank t % world only."""
    n, cin, h, w, cout = 64, 512, 2, 2, 256          # 2 m-tiles x 4 n-tiles = 8 tiles
    fz = _virtual(world)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, cin, h, w, generator=g) * 0.5
    W_full = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    k = cin // world
    a = [cl(x[:, r * k:(r + 1) * k].to(DEV).bfloat16()) for r in range(world)]
    ws = [cl(W_full[:, r * k:(r + 1) * k].to(DEV).bfloat16()) for r in range(world)]
    ref = sum(tb.conv_fwd(a[r].float(), ws[r].float(), 1, 1, False)[0] for r in range(world))
    ops_ = [f.reduce_scatter_conv(0, (n, k, h, w), (cout, k, 3, 3), 1, 1) for f in fz]
    for it in range(3):
        ys = _run_ranks([(lambda r=r: ops_[r](a[r], ws[r])) for r in range(world)])
        # tile (mt, nt): rows = images [32*mt, 32*mt+32) (2x2 maps: 4 pixels per image), columns [64*nt, 64*nt+64)
        for mt in range(2):
            for nt in range(4):
                owner = (nt * 2 + mt) % world
                got = ys[owner][32 * mt:32 * mt + 32, 64 * nt:64 * nt + 64]
                want = ref[32 * mt:32 * mt + 32, 64 * nt:64 * nt + 64]
                assert rel_err(got, want) < 2.5e-2, (mt, nt, owner, it)


@pytest.mark.parametrize("world", [2, 4])
def test_fused_allgather_gemm_virtual_ranks(nb, tb, world):
    """A operand image-sharded over the ranks: the kernel's TMA reads every tile from the owning rank's heap."""
    n, c, h, w, cout = 64, 64, 8, 8, 64
    fz = _virtual(world)
    g = torch.Generator().manual_seed(21)
    x = cl((torch.randn(n, c, h, w, generator=g) * 0.5).to(DEV).bfloat16())
    nl = n // world
    offs = []
    for r, f in enumerate(fz):
        off, buf = f.ag_buffer((nl, c, h, w))
        buf.copy_(x[r * nl:(r + 1) * nl])
        offs.append(off)
    assert len(set(offs)) == 1
    ws = [cl((torch.randn(cout, c, 3, 3, generator=g) / (c * 9) ** 0.5).to(DEV).bfloat16()) for _ in range(world)]
    ops_ = [f.ag_conv(offs[0], (n, c, h, w), (cout, c, 3, 3)) for f in fz]
    for it in range(3):
        ys = _run_ranks([(lambda r=r: ops_[r](None, ws[r])) for r in range(world)])
        for r in range(world):
            ref, _ = tb.conv_fwd(x.float(), ws[r].float(), 1, 1, False)
            assert rel_err(ys[r], ref) < 2e-2


@pytest.mark.parametrize("world", [2, 8])
def test_tp_head_virtual_ranks(nb, world):
    """Column-parallel classifier + softmax-CE + exact backward in one kernel per rank == dense fp32 head."""
    from horizonml_b200.parallel.tp import padded_classes
    n, c, classes = 64, 512, 10
    kpad = padded_classes(classes, world)
    kl = kpad // world
    fz = _virtual(world)
    g = torch.Generator().manual_seed(2)
    feat = cl((torch.randn(n, c, 1, 1, generator=g)).to(DEV).bfloat16())
    Wf = torch.zeros(kpad, c)
    bf = torch.zeros(kpad)
    Wf[:classes] = torch.randn(classes, c, generator=g) / c ** 0.5
    bf[:classes] = torch.randn(classes, generator=g) * 0.1
    labels = torch.randint(0, classes, (n,), generator=g).to(DEV)
    heads = [f.head(n, c, kl) for f in fz]
    Wl = [Wf[r * kl:(r + 1) * kl].contiguous().to(DEV) for r in range(world)]
    bl = [bf[r * kl:(r + 1) * kl].contiguous().to(DEV) for r in range(world)]
    # dense fp32 reference
    pooled = feat.float().view(n, c).requires_grad_(True)
    logits = pooled @ Wf[:classes].to(DEV).t() + bf[:classes].to(DEV)
    loss = torch.nn.functional.cross_entropy(logits, labels)
    Wd = Wf[:classes].to(DEV).clone().requires_grad_(True)
    loss2 = torch.nn.functional.cross_entropy(pooled.detach() @ Wd.t() + bf[:classes].to(DEV), labels)
    loss.backward(); loss2.backward()
    for it in range(3):
        dW = [torch.zeros(kl, c, device=DEV) for _ in range(world)]
        db = [torch.zeros(kl, device=DEV) for _ in range(world)]
        z2 = [torch.zeros(2, device=DEV) for _ in range(world)]
        outs = _run_ranks([(lambda r=r: heads[r](feat, Wl[r], bl[r], labels, 1.0, classes, dW[r], db[r], False, True,
                                                  z2[r])) for r in range(world)])
        for r in range(world):
            lo, correct, dfeat, lg = outs[r]
            assert abs(lo.item() - loss.item()) < 2e-3 * max(1.0, abs(loss.item()))
            assert correct.item() == (logits.argmax(1) == labels).sum().item()
            assert rel_err(lg[:, :classes], logits.detach()) < 1e-3
            assert rel_err(dfeat.view(n, c), pooled.grad) < 1e-2
            lo_k, hi_k = r * kl, min((r + 1) * kl, classes)
            if hi_k > lo_k:
                assert rel_err(dW[r][: hi_k - lo_k], Wd.grad[lo_k:hi_k]) < 1e-3
            assert torch.equal(dfeat, outs[0][2])


@pytest.mark.parametrize("world", [2, 8])
def test_tp_allreduce_bf16_virtual_ranks(nb, world):
    fz = _virtual(world)
    g = torch.Generator().manual_seed(4)
    xs = [(torch.randn(64, 256, 2, 2, generator=g)).to(DEV).bfloat16() for _ in range(world)]
    ref = sum(x.float() for x in xs)
    for it in range(3):
        ys = _run_ranks([(lambda r=r: fz[r].allreduce_bf16(xs[r])) for r in range(world)])
        for r in range(world):
            assert rel_err(ys[r], ref) < 1e-2 and torch.equal(ys[r], ys[0])


def test_tp_model_step_virtual_ranks_matches_dense(tmp_path):
    """Strategy equivalence with the NATIVE kernels on one GPU (tools/tp_virtual_step.py): a tensor-parallel ResNet-18
    (layer3/4 column/row split, column-parallel head, every reduction a fused peer kernel) over W virtual ranks produces
    the dense model's loss and — after reassembling the shards — its gradients, per parameter.  Runs in a subprocess
    with HZ_PDL=0: with programmatic dependent launch the *parked* successor kernels of one virtual rank could occupy
    the SMs another virtual rank's spinning kernel needs (cannot happen with one rank per GPU)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for world in (2, 8):
        out = tmp_path / f"tpv{world}.json"
        env = dict(os.environ, HZ_PDL="0")
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "tp_virtual_step.py"), str(world), str(out)],
                           env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        res = json.load(open(out))
        assert res["library_collectives"] == 0, "a torch.distributed collective ran inside the TP step"
        assert res["native_fallbacks"] == {}, res["native_fallbacks"]
        assert res["loss_rel_err"] < 2e-2
        assert not res["bad"], res["bad"]
