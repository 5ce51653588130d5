"""BatchNorm-backward sums taken by the kernel that produces the BatchNorm's upstream gradient (ops.BNBackLink,
HZ_BN_BWD_IN_DGRAD): the tcgen05 dgrad kernel's kBnBwd instantiation, the max-pool backward and the depthwise dgrad variants
against the sums the separate reduction kernel computes, and whole blocks / the whole ResNet-18 step with the hand-off on
against the same code with it off (launch counts included).  `late` (order 8: new epilogue code in the tcgen05 kernel runs
after the SIMT-only late tests, before the persistent kernel)."""
import pytest
import torch

from test_gpu_blocks import DEV, _run, cl, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=8)]


# ---- BatchNorm-backward sums taken in the consumer's dgrad epilogue (ops.BNBackLink, HZ_BN_BWD_IN_DGRAD) --------------
DGRAD_SHAPES = [  # N, Cin, H, W, Cout, R, stride, pad
    (64, 64, 8, 8, 64, 3, 1, 1), (64, 64, 8, 8, 128, 3, 2, 1), (64, 128, 4, 4, 128, 3, 1, 1), (64, 256, 2, 2, 256, 3, 1, 1),
    (64, 512, 1, 1, 512, 3, 1, 1), (64, 64, 8, 8, 128, 1, 2, 0), (16, 64, 8, 8, 64, 3, 1, 1), (64, 32, 4, 4, 96, 3, 1, 1),
]


@pytest.mark.parametrize("cfg", DGRAD_SHAPES)
@pytest.mark.parametrize("relu,with_addend", [(1, False), (0, False), (1, True), (2, False)])
def test_dgrad_with_bn_backward_sums(cfg, relu, with_addend):
    """conv_dgrad_bnbwd == plain dgrad (bit-identical dx) + the sums the separate channel-reduce kernel would produce."""
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.ops import torch_backend as tb
    N, Cin, H, W, Cout, R, s, p = cfg
    g = torch.Generator().manual_seed(9)
    w = cl((torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(DEV).bfloat16())
    Ho, Wo = (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    dy = cl((torch.randn(N, Cout, Ho, Wo, generator=g) * 0.5).to(DEV).bfloat16())
    add = cl((torch.randn(N, Cin, H, W, generator=g) * 0.5).to(DEV).bfloat16()) if with_addend else None
    y_raw = cl(torch.randn(N, Cin, H, W, generator=g).to(DEV).bfloat16())            # the producing layer's tensors
    mean, invstd = torch.randn(Cin, generator=g).to(DEV) * 0.1, (torch.rand(Cin, generator=g) + 0.5).to(DEV)
    out = cl((torch.randn(N, Cin, H, W, generator=g) * 4).to(DEV).bfloat16())         # ~half <= 0, ~7 % >= 6: both masks
    nb.step_begin(DEV)
    got = nb.conv_dgrad_bnbwd(dy, w, (N, Cin, H, W), s, p, add, out, y_raw, mean, invstd, relu)
    assert got is not None
    dx, sums = got
    dx0 = nb.conv_dgrad(dy, w, (N, Cin, H, W), s, p, add)
    nb.step_end()
    assert torch.equal(dx, dx0)
    # reference sums from the stored bf16 dx, exactly what bn_act_bwd's reduction kernel computes
    mask = 1.0 if relu == 0 else ((out > 0).float() if relu == 1 else ((out > 0) & (out < 6)).float())
    gg = dx.float() * mask
    xhat = (y_raw.float() - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    ref = torch.stack([gg.sum(dim=(0, 2, 3)), (gg * xhat).sum(dim=(0, 2, 3))])
    assert rel_err(sums.view(2, -1), ref) < 1e-3
    # and the apply pass fed with them == the two-kernel BN backward
    gamma = (torch.rand(Cin, generator=g) + 0.5).to(DEV)
    a = nb.bn_act_bwd(dx, out, y_raw, mean, invstd, gamma, relu, False, sums=sums)
    b = nb.bn_act_bwd(dx, out, y_raw, mean, invstd, gamma, relu, False)
    assert rel_err(a[0], b[0]) < 2e-3 and rel_err(a[1], b[1]) < 1e-3 and rel_err(a[2], b[2]) < 1e-3


@pytest.mark.parametrize("cin,cout,stride,hw", [(64, 64, 1, 8), (64, 128, 2, 8), (128, 256, 2, 4), (256, 256, 1, 2),
                                                (512, 512, 1, 1)])
def test_basic_block_with_bn_sums_in_dgrad(cin, cout, stride, hw):
    """HZ_BN_BWD_IN_DGRAD on the native backend: same block output / gradients as with the separate reduction kernel,
    one bn_act_bwd launch less."""
    import horizonml_b200.models.resnet as R
    from horizonml_b200 import ops
    from horizonml_b200.ops import native_backend as nb
    g = torch.Generator().manual_seed(7)
    x0 = cl(torch.randn(64, cin, hw, hw, generator=g).to(DEV).bfloat16())
    dy = cl((torch.randn(64, cout, hw // stride, hw // stride, generator=g) * 0.1).to(DEV).bfloat16())
    res, launches = {}, {}
    try:
        for flag in (False, True):
            R._BN_BWD_IN_DGRAD = flag
            before = nb.LAUNCHES["bn_act_bwd"]
            res[flag] = _run(lambda: R.BasicBlock(cin, cout, stride), x0, dy, "native")
            launches[flag] = nb.LAUNCHES["bn_act_bwd"] - before
    finally:
        R._BN_BWD_IN_DGRAD = False
        ops.set_backend("torch")
    # bn1 <- conv2's dgrad; with a downsample branch also its BN <- bn2's apply kernel
    assert launches[True] == launches[False] - (2 if (stride != 1 or cin != cout) else 1), launches
    (y0, dx0, g0), (y1, dx1, g1) = res[False], res[True]
    assert rel_err(y1, y0) < 1e-2 and rel_err(dx1, dx0) < 2e-2       # (BN statistics: fp32 atomics, last-bit noise)
    for n in g0:
        if g0[n].abs().max().item() > 1e-6:
            assert rel_err(g1[n], g0[n]) < 2e-2, (n, rel_err(g1[n], g0[n]))


def test_resnet18_step_with_bn_sums_in_dgrad():
    """Whole model: 19 of the 20 BatchNorm-backward reduction kernels are gone (bn1 of every block and bn2 of every block
    but the last take their sums from a dgrad epilogue, the stem's bn1 from the max-pool backward kernel, the three
    downsample BNs from their block's bn2 apply kernel), the first-step loss is unchanged, gradients agree with the
    two-kernel path to the run-to-run noise of the fp32 atomics, and training still makes progress."""
    import horizonml_b200.models.resnet as R
    from horizonml_b200 import ops
    from horizonml_b200.models.flat import FlatAdam, FlatParams
    from horizonml_b200.ops import native_backend as nb
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(DEV)
    labels = torch.randint(0, 10, (64,), generator=g).to(DEV)
    res = {}
    try:
        ops.set_backend("native")
        for flag in (False, True):
            R._BN_BWD_IN_DGRAD = flag
            model = R.resnet18(10, seed=0).to(DEV).train()
            flat = FlatParams(list(model.named_parameters()), DEV, torch.bfloat16)
            opt = FlatAdam(flat, lr=1e-3)
            losses = []
            for it in range(3):
                x = ops.stem_prepare(images.permute(0, 3, 1, 2), dtype=torch.bfloat16)
                ops.step_begin(DEV)
                flat.begin_step()
                before = nb.LAUNCHES["bn_act_bwd"]
                loss, _ = model.forward_loss(x, labels)
                loss.backward()
                ops.join_side()
                ops.step_end()
                if it == 0:
                    res[flag] = (flat.grad.clone(), nb.LAUNCHES["bn_act_bwd"] - before)
                opt.step()
                losses.append(float(loss.detach()))
            res[("loss", flag)] = losses
    finally:
        R._BN_BWD_IN_DGRAD = False
        ops.set_backend("torch")
    # 15 dgrad hand-offs + the pool's + 3 downsample BNs
    assert res[False][1] == 40 and res[True][1] == 21, (res[False][1], res[True][1])
    l0, l1 = res[("loss", False)], res[("loss", True)]
    assert abs(l0[0] - l1[0]) < 1e-3 and all(v == v for v in l1) and l1[-1] < l1[0], (l0, l1)   # (loss: fp32 atomics)
    cos = torch.nn.functional.cosine_similarity(res[False][0].flatten(), res[True][0].flatten(), dim=0).item()
    assert cos > 0.9, cos


@pytest.mark.parametrize("N,C,H", [(64, 128, 4), (64, 256, 2), (64, 512, 1), (64, 64, 8), (5, 128, 3)])
@pytest.mark.parametrize("relu,own_ready", [(1, False), (1, True), (0, False)])
def test_bn_backward_apply_with_residual_bn_sums(N, C, H, relu, own_ready):
    """bn_act_bwd_res == bn_act_bwd (dy, dres, dgamma, dbeta) + the backward sums of the BatchNorm that produced the
    residual, taken from the stored dres; with and without this layer's own sums already final."""
    from horizonml_b200.ops import native_backend as nb
    from horizonml_b200.ops import torch_backend as tb
    g = torch.Generator().manual_seed(21)
    mk = lambda scale=1.0: cl((torch.randn(N, C, H, H, generator=g) * scale).to(DEV).bfloat16())      # noqa: E731
    dout, out, y_raw, res_yraw = mk(0.5), mk(), mk(), mk()
    vec = lambda: (torch.randn(C, generator=g).to(DEV) * 0.1, (torch.rand(C, generator=g) + 0.5).to(DEV))   # noqa: E731
    (mean, invstd), (rmean, rinvstd) = vec(), vec()
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV)
    slots = lambda: (tb.GradSlot(torch.zeros(C, device=DEV), False), tb.GradSlot(torch.zeros(C, device=DEV), False))  # noqa: E731
    s0, s1 = slots(), slots()
    dy0, _, _, dres0 = nb.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, relu, True, *s0)
    own = None
    if own_ready:
        gg = dout.float() * ((out > 0).float() if relu else 1.0)
        xh = (y_raw.float() - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
        own = torch.stack([gg.sum(dim=(0, 2, 3)), (gg * xh).sum(dim=(0, 2, 3))]).contiguous()
    nb.step_begin(DEV)
    got = nb.bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma, relu, *s1, own, res_yraw, rmean, rinvstd)
    nb.step_end()
    assert got is not None
    dy1, dres1, rs = got
    assert torch.equal(dres1, dres0)
    assert rel_err(dy1, dy0) < 2e-3 and rel_err(s1[0].t, s0[0].t) < 1e-3 and rel_err(s1[1].t, s0[1].t) < 1e-3
    gr = dres0.float()
    xhat = (res_yraw.float() - rmean.view(1, -1, 1, 1)) * rinvstd.view(1, -1, 1, 1)
    ref = torch.stack([gr.sum(dim=(0, 2, 3)), (gr * xhat).sum(dim=(0, 2, 3))])
    assert rel_err(rs.view(2, -1), ref) < 1e-3
    # outside a step (no pre-zeroed arena): the launcher zeroes its own scratch
    got2 = nb.bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma, relu, *slots(), own, res_yraw, rmean, rinvstd)
    assert rel_err(got2[2].view(2, -1), ref) < 1e-3 and rel_err(got2[0], dy0) < 2e-3


@pytest.mark.parametrize("relu", [1, 0])
def test_maxpool_backward_with_bn_sums(relu):
    """maxpool_bwd_bn == max-pool backward (bit-identical dx) + the sums the reduction kernel would produce."""
    from horizonml_b200.ops import native_backend as nb
    g = torch.Generator().manual_seed(4)
    x = cl(torch.randn(64, 64, 16, 16, generator=g).to(DEV).bfloat16())                 # = the stem's BN output
    y_raw = cl(torch.randn(64, 64, 16, 16, generator=g).to(DEV).bfloat16())
    mean, invstd = torch.randn(64, generator=g).to(DEV) * 0.1, (torch.rand(64, generator=g) + 0.5).to(DEV)
    y, aux = nb.maxpool_fwd(x, True)
    dy = cl(torch.randn(64, 64, 8, 8, generator=g).to(DEV).bfloat16())
    nb.step_begin(DEV)
    got = nb.maxpool_bwd_bn(dy, aux, x, y_raw, mean, invstd, relu)
    nb.step_end()
    assert got is not None
    dx, sums = got
    assert torch.equal(dx, nb.maxpool_bwd(dy, aux))
    gg = dx.float() * ((x > 0).float() if relu else 1.0)
    xhat = (y_raw.float() - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    ref = torch.stack([gg.sum(dim=(0, 2, 3)), (gg * xhat).sum(dim=(0, 2, 3))])
    assert rel_err(sums.view(2, -1), ref) < 1e-3


@pytest.mark.parametrize("inp,oup,stride,t,hw", [(24, 24, 1, 6, 8), (24, 32, 2, 6, 8), (96, 160, 2, 6, 2)])
def test_mobilenet_blocks_with_bn_sums_in_dgrad(inp, oup, stride, t, hw):
    """A stack of two inverted-residual blocks with HZ_BN_BWD_IN_DGRAD on the native backend: the depthwise BatchNorms (ReLU6
    mask, channel counts that are not powers of two) get their sums from the project convs' dgrad kernels, the expand
    BatchNorms from the depthwise convs' dgrad kernels, the first block's project BN from the second block's expand conv
    (skip share folded in): fewer bn_act_bwd launches, same gradients as the separate reduction kernels."""
    import horizonml_b200.models.resnet as R
    from horizonml_b200 import ops
    from horizonml_b200.models.mobilenet import InvertedResidual
    from horizonml_b200.ops import native_backend as nb
    g = torch.Generator().manual_seed(7)
    x0 = cl(torch.randn(64, inp, hw, hw, generator=g).to(DEV).bfloat16())
    ho = (hw - 1) // stride + 1
    dy = cl((torch.randn(64, oup, ho, ho, generator=g) * 0.1).to(DEV).bfloat16())
    mk = lambda: torch.nn.Sequential(InvertedResidual(inp, oup, stride, t), InvertedResidual(oup, oup, 1, t))     # noqa: E731
    res, launches = {}, {}
    try:
        for flag in (False, True):
            R._BN_BWD_IN_DGRAD = flag
            before = nb.LAUNCHES["bn_act_bwd"]
            res[flag] = _run(mk, x0, dy, "native")
            launches[flag] = nb.LAUNCHES["bn_act_bwd"] - before
    finally:
        R._BN_BWD_IN_DGRAD = False
        ops.set_backend("torch")
    assert launches[True] == launches[False] - 5, launches          # 2 expand BNs + 2 depthwise BNs + the first block's project BN
    (y0, dx0, g0), (y1, dx1, g1) = res[False], res[True]
    assert rel_err(y1, y0) < 1e-2 and rel_err(dx1, dx0) < 3e-2
    for n in g0:
        if g0[n].abs().max().item() > 1e-6:
            assert rel_err(g1[n], g0[n]) < 3e-2, (n, rel_err(g1[n], g0[n]))


@pytest.mark.parametrize("N,C,H,stride", [(64, 96, 16, 2), (64, 144, 8, 1), (64, 576, 2, 1), (64, 960, 1, 1), (3, 24, 7, 2)])
@pytest.mark.parametrize("act", [2, 1, 0])
def test_depthwise_dgrad_with_bn_backward_sums(N, C, H, stride, act):
    """dwconv_dgrad_bnbwd == depthwise dgrad (bit-identical dx) + the producing BatchNorm's backward sums (ReLU6 mask)."""
    from horizonml_b200.ops import native_backend as nb
    g = torch.Generator().manual_seed(12)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(DEV).bfloat16()
    Ho = (H - 1) // stride + 1
    dy = cl((torch.randn(N, C, Ho, Ho, generator=g) * 0.5).to(DEV).bfloat16())
    y_raw = cl(torch.randn(N, C, H, H, generator=g).to(DEV).bfloat16())
    out = cl((torch.randn(N, C, H, H, generator=g) * 4).to(DEV).bfloat16())
    mean, invstd = torch.randn(C, generator=g).to(DEV) * 0.1, (torch.rand(C, generator=g) + 0.5).to(DEV)
    nb.step_begin(DEV)
    got = nb.dwconv_dgrad_bnbwd(dy, w, (N, C, H, H), stride, out, y_raw, mean, invstd, act)
    nb.step_end()
    assert got is not None
    dx, sums = got
    assert torch.equal(dx, nb.dwconv_dgrad(dy, w, (N, C, H, H), stride))
    mask = 1.0 if act == 0 else ((out > 0).float() if act == 1 else ((out > 0) & (out < 6)).float())
    gg = dx.float() * mask
    xhat = (y_raw.float() - mean.view(1, -1, 1, 1)) * invstd.view(1, -1, 1, 1)
    ref = torch.stack([gg.sum(dim=(0, 2, 3)), (gg * xhat).sum(dim=(0, 2, 3))])
    assert rel_err(sums.view(2, -1), ref) < 1e-3
