"""CPU tier for the MobileNetV2 op set: the depthwise-conv autograd op and the ReLU6 activation code on the PyTorch-op
backend against plain torch autograd, the host build of the depthwise kernels' index logic, and the pybind signatures
of the bindings the native backend calls for them (a wrong arity would otherwise only show on a GPU box)."""
import os
import shutil
import subprocess

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("stride,hw", [(1, 8), (2, 8), (2, 7), (1, 1)])
def test_dwconv_bn_act_matches_autograd(stride, hw):
    from horizonml_b200 import ops
    ops.set_backend("torch")
    g = torch.Generator().manual_seed(stride * 10 + hw)
    C = 24
    x = torch.randn(6, C, hw, hw, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) * 3 + 1).requires_grad_(True)      # some outputs beyond the ReLU6 cap
    beta = (torch.randn(C, generator=g) + 2).requires_grad_(True)
    rm, rv = torch.zeros(C), torch.ones(C)
    out = ops.dwconv_bn_act(x, w, gamma, beta, rm, rv, stride=stride, act=2, training=True)
    dout = torch.randn(out.shape, generator=g)
    out.backward(dout)
    got = [x.grad.clone(), w.grad.clone(), gamma.grad.clone(), beta.grad.clone()]
    for t in (x, w, gamma, beta):
        t.grad = None
    rm2, rv2 = torch.zeros(C), torch.ones(C)
    ref = F.relu6(F.batch_norm(F.conv2d(x, w, None, stride, 1, 1, C), rm2, rv2, gamma, beta, True, 0.1, 1e-5))
    ref.backward(dout)
    assert torch.allclose(out, ref, atol=1e-5) and (ref == 6).float().mean() > 0.01
    assert torch.allclose(rm, rm2, atol=1e-6) and torch.allclose(rv, rv2, atol=1e-5)
    for a, b, name in zip(got, (x.grad, w.grad, gamma.grad, beta.grad), ("dx", "dw", "dgamma", "dbeta")):
        assert torch.allclose(a, b, atol=2e-4, rtol=1e-4), (name, (a - b).abs().max().item())


def test_conv_bn_act_relu6_code_matches_autograd():
    from horizonml_b200 import ops
    ops.set_backend("torch")
    g = torch.Generator().manual_seed(2)
    x = torch.randn(4, 16, 6, 6, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(24, 16, 1, 1, generator=g) / 4).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gamma = (torch.rand(24, generator=g) * 3 + 1).requires_grad_(True)
    beta = (torch.randn(24, generator=g) + 2).requires_grad_(True)
    out = ops.conv_bn_act(x, w, gamma, beta, torch.zeros(24), torch.ones(24), stride=1, pad=0, relu=2)
    dout = torch.randn(out.shape, generator=g)
    out.backward(dout)
    gx, gw = x.grad.clone(), w.grad.clone()
    x.grad = w.grad = None
    ref = F.relu6(F.batch_norm(F.conv2d(x, w), torch.zeros(24), torch.ones(24), gamma, beta, True, 0.1, 1e-5))
    ref.backward(dout)
    assert torch.allclose(out, ref, atol=1e-5) and (ref == 6).any()
    assert torch.allclose(gx, x.grad, atol=2e-4) and torch.allclose(gw, w.grad, atol=2e-4)


def test_host_depthwise_logic(tmp_path):
    """csrc/tests/dw_host_test.cu: the depthwise kernels' thread layout, row walk and forward / dgrad / wgrad index math
    executed on the host, thread by thread, against the convolution definition (no GPU, no kernel launch)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "dw_host_test")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                        os.path.join(ROOT, "csrc"), "-o", exe, os.path.join(ROOT, "csrc", "tests", "dw_host_test.cu")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host depthwise logic ok" in r.stdout, r.stdout + r.stderr


def test_native_binding_signatures_for_mobilenet_ops():
    """Call the new / changed bindings with CPU tensors: the device check must fire (RuntimeError), i.e. pybind accepted
    the argument list the native backend passes."""
    from horizonml_b200.ops import _ext
    C = _ext.load(required=False)
    if C is None:
        pytest.skip("extension not built")
    x = torch.zeros(2, 16, 4, 4, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.zeros(16, 1, 3, 3, dtype=torch.bfloat16)
    f = torch.zeros(16)
    calls = {
        "dwconv_fwd": lambda: C.dwconv_fwd(x, w, 1, True, None),
        "dwconv_dgrad": lambda: C.dwconv_dgrad(x, w, [2, 16, 4, 4], 1),
        "dwconv_wgrad": lambda: C.dwconv_wgrad(x, x, torch.zeros(16, 1, 3, 3), 1, False, True),
        "dwconv_dgrad_bnbwd": lambda: C.dwconv_dgrad_bnbwd(x, w, [2, 16, 4, 4], 1, x, x, f, f, None, True),
        "bn_act_fwd": lambda: C.bn_act_fwd(x, torch.zeros(2, 16), f, f, f, f, 0.1, 1e-5, None, 2, True),
        "bn_act_bwd": lambda: C.bn_act_bwd(x, x, x, f, f, f, 2, False, f, f, False, False, None, False),
        "bn_act_bwd(sums)": lambda: C.bn_act_bwd(x, x, x, f, f, f, 1, False, f, f, False, False, torch.zeros(2, 16), True),
        "bn_act_bwd_res": lambda: C.bn_act_bwd_res(x, x, x, f, f, f, 1, f, f, False, False, torch.zeros(2, 16), True,
                                                   x, f, f, torch.zeros(2, 16)),
        "bn_act_bwd_res(no scratch)": lambda: C.bn_act_bwd_res(x, x, x, f, f, f, 0, f, f, True, True, None, False,
                                                               x, f, f, None),
        "conv_dgrad_bnbwd": lambda: C.conv_dgrad_bnbwd(x, torch.zeros(16, 16, 3, 3, dtype=torch.bfloat16).contiguous(
            memory_format=torch.channels_last), [2, 16, 4, 4], 1, 1, None, False, x, x, f, f, None, False),
    }
    for name, fn in calls.items():
        with pytest.raises(RuntimeError, match="CUDA tensor"):
            fn()
    assert C.dwconv_ok(64, 16, 16, 96, 2) and not C.dwconv_ok(64, 16, 16, 12, 1) and not C.dwconv_ok(64, 16, 16, 16, 3)
    # 1: default BatchNorm instantiation (C/8 a power of two), 2: generic one, 0: not a multiple of 8
    assert [C.channel_ok(c) for c in (64, 24, 1280, 12)] == [1, 2, 2, 0]


def test_mobilenet_bn_backward_sums_hand_off_matches_plain_backward():
    """HZ_BN_BWD_IN_DGRAD on MobileNetV2 (PyTorch-op oracle of the fused dgrad): every depthwise BN gets its backward sums
    from the project conv's dgrad, every project BN from the next expand conv's (or the head conv's) dgrad, every expand
    BN from the depthwise conv's dgrad — 17 + 17 + 16 hand-offs (50 of 52 BatchNorms), ReLU6 masks and skip connections
    included — and loss / gradients do not change."""
    import horizonml_b200.models.resnet as R
    from horizonml_b200 import ops
    from horizonml_b200.models.mobilenet import mobilenet_v2
    from horizonml_b200.ops import torch_backend as tb
    ops.set_backend("torch")
    calls = {"n": 0}
    orig = tb.conv_dgrad_bnbwd

    def counted(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    tb.conv_dgrad_bnbwd = counted
    dw_calls = {"n": 0}
    orig_dw = tb.dwconv_dgrad_bnbwd

    def counted_dw(*a, **k):
        dw_calls["n"] += 1
        return orig_dw(*a, **k)
    tb.dwconv_dgrad_bnbwd = counted_dw
    try:
        xb = torch.randn(8, 3, 32, 32, generator=torch.Generator().manual_seed(4)).contiguous(memory_format=torch.channels_last)
        yb = torch.randint(0, 10, (8,), generator=torch.Generator().manual_seed(5))
        res = []
        for flag in (False, True):
            R._BN_BWD_IN_DGRAD = flag
            model = mobilenet_v2(10, seed=0, dropout=0.0).train()
            loss, _ = model.forward_loss(xb, yb)
            loss.backward()
            res.append([loss.detach()] + [p.grad for p in model.parameters()])
        for a, b in zip(*res):
            assert torch.allclose(a, b, atol=5e-5, rtol=1e-3), (a - b).abs().max()
        assert calls["n"] == 34 and dw_calls["n"] == 16, (calls, dw_calls)      # + the 16 expand BNs <- depthwise dgrad
    finally:
        tb.conv_dgrad_bnbwd = orig
        tb.dwconv_dgrad_bnbwd = orig_dw
        R._BN_BWD_IN_DGRAD = False
