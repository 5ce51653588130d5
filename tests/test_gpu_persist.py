"""The persistent (throughput) variant of the tcgen05 convolution kernel — csrc/conv_gemm.cu ``igemm_persist_kernel``:
one CTA per SM walking a static tile schedule, TMA producer running ahead across tiles, two TMEM accumulators, epilogue
warps overlapped with the next tile's MMAs — against the PyTorch fp32 oracle and against the one-tile-per-CTA kernel
(same operands, same K order: identical bf16 outputs whenever that kernel runs without split-K).

Opt-in (``HZ_CONV_PERSIST`` / ``C.conv_set_persist``) and `late`: written after the round's GPU budget was spent, this
file is the kernel's first execution on hardware.  It is collected last of all (order=9): new tcgen05 code is the
least certain item of the late tier."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.late(order=9)]
DEV = "cuda:0"


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-6)).item()


@pytest.fixture(scope="module")
def nb():
    from horizonml_b200.ops import native_backend
    return native_backend


@pytest.fixture(scope="module")
def tb():
    from horizonml_b200.ops import torch_backend
    return torch_backend


@pytest.fixture(params=[1, 2], ids=["wide", "n64"])
def persist(nb, request):
    """1: 128-column tiles where the channel count is a multiple of 128 (else 64), 2: 64-column tiles everywhere."""
    old = nb.C.conv_set_persist(request.param)
    yield request.param
    nb.C.conv_set_persist(old)


# N, Cin, H, W, Cout, R, stride, pad
SHAPES = [
    (64, 64, 8, 8, 64, 3, 1, 1),          # layer1: 32 tiles, one per CTA (persistent loop runs once)
    (1024, 64, 8, 8, 64, 3, 1, 1),        # layer1 at batch 1024: 512 tiles on 148 CTAs -> 3-4 tiles per CTA, ring wraps
    (512, 64, 8, 8, 128, 3, 2, 1),        # stride-2 forward (parity views), stride-2 dgrad (4 classes)
    (512, 64, 8, 8, 128, 1, 2, 0),        # 1x1 stride 2: dgrad classes without taps (zero tiles, no accumulator)
    (256, 128, 4, 4, 256, 3, 1, 1),       # 2 k-blocks per tap, 4 n-tiles
    (512, 256, 2, 2, 512, 3, 1, 1),       # dead taps, 8 n-tiles
    (64, 512, 1, 1, 512, 3, 1, 1),        # centre tap only, one m-tile with rows past the last image
    (300, 32, 4, 4, 96, 3, 1, 1),         # narrow / non-multiple-of-64 channels, last tile partially past the batch
    (16384, 64, 1, 1, 32, 1, 1, 0),       # the stem GEMM shape of MobileNetV2 (im2col rows x 64 -> 32)
]


def _data(cfg, seed=1):
    N, Cin, H, W, Cout, R, s, p = cfg
    g = torch.Generator().manual_seed(seed)
    x = cl((torch.randn(N, Cin, H, W, generator=g) * 0.5).to(DEV).bfloat16())
    w = cl((torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(DEV).bfloat16())
    Ho, Wo = (H + 2 * p - R) // s + 1, (W + 2 * p - R) // s + 1
    dy = cl((torch.randn(N, Cout, Ho, Wo, generator=g) * 0.5).to(DEV).bfloat16())
    add = cl((torch.randn(N, Cin, H, W, generator=g) * 0.5).to(DEV).bfloat16())
    return x, w, dy, add


@pytest.mark.parametrize("cfg", SHAPES)
def test_persistent_forward(nb, tb, persist, cfg):
    x, w, _, _ = _data(cfg)
    s, p = cfg[6], cfg[7]
    assert nb._conv_ok(x.shape, w.shape, s, p)
    y, stats = nb.conv_fwd(x, w, s, p, True)
    yr, sr = tb.conv_fwd(x.float(), w.float(), s, p, True)
    assert rel_err(y, yr) < 2e-2 and rel_err(stats, sr) < 2e-2
    y2, none = nb.conv_fwd(x, w, s, p, False)
    assert none is None and torch.equal(y, y2)                     # deterministic, with and without the BN sums
    nb.C.conv_set_persist(0)                                       # the hardware-verified kernel on the same operands
    y0, st0 = nb.conv_fwd(x, w, s, p, True)
    nb.C.conv_set_persist(persist)
    assert rel_err(y, y0) < 8e-3 and rel_err(stats, st0) < 5e-3    # (cluster split-K there may reorder the fp32 sum: 1 bf16 ulp)


@pytest.mark.parametrize("cfg", SHAPES)
@pytest.mark.parametrize("with_addend", [False, True])
def test_persistent_dgrad(nb, tb, persist, cfg, with_addend):
    x, w, dy, add = _data(cfg, seed=2)
    s, p = cfg[6], cfg[7]
    a = add if with_addend else None
    dx = nb.conv_dgrad(dy, w, x.shape, s, p, a)
    ref = tb.conv_dgrad(dy.float(), w.float(), x.shape, s, p, a.float() if a is not None else None)
    assert rel_err(dx, ref) < 2e-2
    nb.C.conv_set_persist(0)
    dx0 = nb.conv_dgrad(dy, w, x.shape, s, p, a)
    nb.C.conv_set_persist(persist)
    assert rel_err(dx, dx0) < 8e-3


@pytest.mark.parametrize("cfg", [SHAPES[1], SHAPES[2], SHAPES[8]])
def test_throughput_mode_wgrad_splits(nb, tb, persist, cfg):
    """In throughput mode a long pixel reduction (>= 512 k-blocks of 64 pixels) is split over two waves of CTAs."""
    x, w, dy, _ = _data(cfg, seed=3)
    s, p = cfg[6], cfg[7]
    Cout, Cin, R = cfg[4], cfg[1], cfg[5]
    gv = torch.zeros(Cout, R, R, Cin, device=DEV).permute(0, 3, 1, 2)          # storage [Cout, R, S, Cin]
    ref = torch.zeros(Cout, Cin, R, R, device=DEV)
    nb.conv_wgrad(dy, x, w.shape, s, p, gv, False)
    tb.conv_wgrad(dy.float(), x.float(), w.shape, s, p, ref, False)
    assert rel_err(gv, ref) < 2e-2


def test_persistent_kernel_in_a_training_step(nb, persist):
    """ResNet-18 forward + backward + Adam at batch 64 with every forward / dgrad convolution on the persistent kernel:
    loss and gradient must match the default kernels' (same seed, same data), and training must make progress."""
    from horizonml_b200 import ops
    from horizonml_b200.models.flat import FlatAdam, FlatParams
    from horizonml_b200.models.resnet import resnet18
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (64, 32, 32, 3), dtype=torch.uint8, generator=g).to(DEV)
    labels = torch.randint(0, 10, (64,), generator=g).to(DEV)
    res = {}
    try:
        ops.set_backend("native")
        for mode in (persist, 0):
            nb.C.conv_set_persist(mode)
            model = resnet18(10, seed=0).to(DEV).train()
            flat = FlatParams(list(model.named_parameters()), DEV, torch.bfloat16)
            opt = FlatAdam(flat, lr=1e-3)
            losses = []
            for it in range(3):
                xb = ops.stem_prepare(images.permute(0, 3, 1, 2), dtype=torch.bfloat16)
                flat.begin_step()
                loss, _ = model.forward_loss(xb, labels)
                loss.backward()
                ops.join_side()
                if it == 0:
                    res[mode] = flat.grad.clone()
                opt.step()
                losses.append(float(loss.detach()))
            res[("loss", mode)] = losses
    finally:
        ops.set_backend("torch")
    l1, l0 = res[("loss", persist)], res[("loss", 0)]
    assert all(v == v for v in l1) and abs(l1[0] - l0[0]) < 2e-2 * max(1.0, abs(l0[0])) and l1[-1] < l1[0]
    cos = torch.nn.functional.cosine_similarity(res[persist].flatten(), res[0].flatten(), dim=0).item()
    assert cos > 0.9, cos          # (two runs of the SAME kernels differ by cos 0.95-0.99: fp32 atomics reorder)
