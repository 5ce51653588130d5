"""Per-block gradient check of the dense model on the native kernels ("per-layer gradient comparison with a
tight tolerance instead of a global cosine"): every distinct BasicBlock of ResNet-18 at CIFAR shapes — and every distinct
inverted-residual block of MobileNetV2 — runs forward + backward ONCE from the same input and the same upstream gradient
on the native backend and on the PyTorch-op oracle.  One block deep there is no chaotic amplification through a stack of
BatchNorm layers, so every tensor (output, input gradient, each parameter gradient) has to agree to bf16 accuracy: a
dropped filter tap, a wrong stride-2 parity class or a mis-fused residual gradient in any single layer fails here even
though the whole-model cosine test would still pass.  `late`: written after the round's GPU budget was spent."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=1)]
DEV = "cuda:0"


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


def _run(block_fn, x0, dy, backend):
    """fresh block (same seed) on `backend`: output, dx and parameter gradients (flat store, fp32)"""
    from horizonml_b200 import ops
    from horizonml_b200.models.flat import FlatParams
    ops.set_backend(backend)
    torch.manual_seed(0)
    blk = block_fn().to(DEV).train()
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if p.dim() == 1 and n.endswith("weight"):
                p.uniform_(0.5, 1.5)                      # BN gammas away from 1 so that their gradients matter
    flat = FlatParams(list(blk.named_parameters()), DEV, torch.bfloat16)
    x = x0.clone().requires_grad_(True)
    ops.step_begin(DEV)
    flat.begin_step()
    y = blk(x)
    y.backward(dy)
    ops.join_side()
    ops.step_end()
    torch.cuda.synchronize()
    grads = {n: p.main_grad.detach().float().clone() for n, p in blk.named_parameters()}
    return y.detach().float(), x.grad.detach().float(), grads


def _compare(block_fn, cin, hw, cout_hw):
    from horizonml_b200 import ops
    from horizonml_b200.ops import native_backend as nb
    g = torch.Generator().manual_seed(7)
    x0 = cl(torch.randn(64, cin, hw, hw, generator=g).to(DEV).bfloat16())
    cout, ho = cout_hw
    dy = cl((torch.randn(64, cout, ho, ho, generator=g) * 0.1).to(DEV).bfloat16())
    before = sum(nb.FALLBACKS.values())
    try:
        yn, dxn, gn = _run(block_fn, x0, dy, "native")
        assert sum(nb.FALLBACKS.values()) == before, dict(nb.FALLBACKS)
        yo, dxo, go = _run(block_fn, x0, dy, "torch")
    finally:
        ops.set_backend("torch")
    assert rel_err(yn, yo) < 3e-2, ("y", rel_err(yn, yo))
    # dx: a ReLU mask that flips where the two backends round a pre-activation to different sides of zero moves ONE element
    # by a whole upstream-gradient value (the residual path adds it unfiltered) — norm-relative error for the tensor, the
    # elementwise bound only guards against gross errors
    nrm = ((dxn - dxo).norm() / (dxo.norm() + 1e-12)).item()
    assert nrm < 2e-2 and rel_err(dxn, dxo) < 0.5, ("dx", nrm, rel_err(dxn, dxo))
    for n in go:
        if go[n].abs().max().item() < 1e-6:
            continue
        # bf16 activations: elementwise agreement to a few percent of the tensor's scale and near-perfect alignment
        e = rel_err(gn[n], go[n])
        c = torch.nn.functional.cosine_similarity(gn[n].flatten(), go[n].flatten(), dim=0).item()
        assert e < 8e-2 and c > 0.995, (n, e, c)


@pytest.mark.parametrize("cin,cout,stride,hw", [(64, 64, 1, 8), (64, 128, 2, 8), (128, 128, 1, 4), (128, 256, 2, 4),
                                                (256, 256, 1, 2), (256, 512, 2, 2), (512, 512, 1, 1)])
def test_resnet_basic_block_native_vs_oracle(cin, cout, stride, hw):
    from horizonml_b200.models.resnet import BasicBlock
    _compare(lambda: BasicBlock(cin, cout, stride), cin, hw, (cout, hw // stride))


@pytest.mark.parametrize("inp,oup,stride,t,hw", [(32, 16, 1, 1, 16), (16, 24, 2, 6, 16), (24, 24, 1, 6, 8), (24, 32, 2, 6, 8),
                                                 (32, 32, 1, 6, 4), (32, 64, 2, 6, 4), (64, 96, 1, 6, 2), (96, 160, 2, 6, 2),
                                                 (160, 160, 1, 6, 1), (160, 320, 1, 6, 1)])
def test_mobilenet_inverted_residual_native_vs_oracle(inp, oup, stride, t, hw):
    from horizonml_b200.models.mobilenet import InvertedResidual
    _compare(lambda: InvertedResidual(inp, oup, stride, t), inp, hw, (oup, (hw - 1) // stride + 1))


@pytest.mark.parametrize("n", [640, 1024, 2048, 5000])        # 5000: three launches of <= 2048 samples
def test_head_beyond_48k_of_shared_memory(n):
    """The classifier head's weight-gradient kernel stages dlogits[N][16] in shared memory: above N = 640 it needs the
    opt-in dynamic shared-memory size (a --batch_size 1024 step failed at this launch before)."""
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.ops import torch_backend as tb
    g = torch.Generator().manual_seed(5)
    f = cl(torch.randn(n, 512, 1, 1, generator=g).to(DEV).bfloat16())
    W = (torch.randn(16, 512, generator=g) * 0.05).to(DEV)
    b = (torch.randn(16, generator=g) * 0.1).to(DEV)
    lab = torch.randint(0, 10, (n,), generator=g).to(DEV)
    dW, db = torch.zeros(16, 512, device=DEV), torch.zeros(16, device=DEV)
    dW2, db2 = torch.zeros_like(dW), torch.zeros_like(db)
    l, c, df, lg = nb.head_fwd_bwd(f, W, b, lab, 1.0, 10, dW, db, False, True)
    l2, c2, df2, lg2 = tb.head_fwd_bwd(f, W, b, lab, 1.0, 10, dW2, db2, False, True)
    torch.cuda.synchronize()
    assert abs(l.item() - l2.item()) < 1e-3 * max(1, abs(l2.item())) and c.item() == c2.item()
    assert rel_err(df, df2) < 1e-2 and rel_err(dW, dW2) < 1e-3 and rel_err(db, db2) < 1e-3
