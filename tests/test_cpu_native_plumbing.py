"""The native backend's Python layer, driven on CPU tensors against a shim of the CUDA extension.

No kernel can run here, but everything between the model and the kernel launch can: ``ops/functional.py`` →
``ops/native_backend.py`` → the pybind argument lists of ``csrc/bindings.cpp``.  Every call the native backend makes is
(1) forwarded to the REAL binding with the CPU tensors — pybind must accept the argument list (no TypeError) and the
binding must then stop at its device check — and (2) answered by an emulation of the kernel built from the PyTorch-op
oracle, honouring the binding's buffer contract (statistics accumulated into the pre-zeroed arena slice, gradients
written / accumulated in place into the flat bucket views).  A whole ResNet-18 / MobileNetV2 training step through that
path must then reproduce the PyTorch-op backend's loss and gradients with no fallback: a mis-ordered argument, a
missing return value, a wrong activation code, a hand-off link feeding the wrong layer or a stale arena slice shows up
here instead of on the first GPU run (native_backend.py was edited after the round's last GPU access)."""
import os
from collections import Counter

import pytest
import torch
import torch.nn.functional as F

from horizonml_b200.ops import _ext
from horizonml_b200.ops import torch_backend as tb

BF16 = torch.bfloat16


class ShimC:
    """Stands in for horizonml_b200._C: real pybind signature check + oracle emulation per binding."""

    def __init__(self, real):
        self._real = real
        self.calls = Counter()

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name == "conv_supported":             # the real one also asks for the CUDA driver's tensor-map encoder
            return self._real.conv_shape_ok
        if name == "PeerComm":                   # peer-memory communicator: emulated ranks of one process (see FakePeerComm)
            FakePeerComm._real = self._real.PeerComm
            return FakePeerComm
        real = getattr(self._real, name)         # AttributeError = the backend calls a binding that does not exist
        emu = getattr(self, "_emu_" + name, None)

        def call(*args):
            self.calls[name] += 1
            try:
                out = real(*args)
            except TypeError as e:                # pybind rejected the argument list
                raise AssertionError(f"C.{name}: argument list rejected by the binding: {str(e)[:400]}") from None
            except RuntimeError as e:
                msg = str(e)
                assert "CUDA tensor" in msg or "is_cuda" in msg or "non-CUDA DeviceType" in msg, \
                    f"C.{name} failed before its device check: {msg[:300]}"
                assert emu is not None, f"no emulation for C.{name}"
                return emu(*args)
            assert emu is None, f"C.{name} accepted CPU tensors"      # pure host helpers only (channel_ok, ...)
            return out
        return call

    # ---- convolution
    @staticmethod
    def _stats_out(sums, pre):
        if pre is None:
            return sums
        pre.view(2, -1).add_(sums)                # the kernels accumulate into the pre-zeroed arena slice
        return pre

    def _emu_conv_fwd(self, x, w, stride, pad, want_stats, pre, stable):
        y, sums = tb.conv_fwd(x, w, stride, pad, want_stats)
        return y, (self._stats_out(sums, pre) if want_stats else None)

    @staticmethod
    def _im2col(x, r, stride, pad, kp):
        cols = F.unfold(x.float(), r, padding=pad, stride=stride)                 # [N, Cin*r*r, L], K order (cin, r, s)
        n, k, L = cols.shape
        cin = x.shape[1]
        A = cols.view(n, cin, r, r, L).permute(0, 4, 2, 3, 1).reshape(n * L, r * r * cin)   # K order (r, s, cin)
        return F.pad(A, (0, kp - k)).to(x.dtype).contiguous()

    def _emu_stem_pack(self, x, w2d, r, stride, pad, kp):
        return self._im2col(x, r, stride, pad, kp), F.pad(w2d, (0, kp - w2d.shape[1])).contiguous()

    def _emu_im2col_small(self, x, r, stride, pad, kp):
        return self._im2col(x, r, stride, pad, kp)

    def _emu_conv_dgrad(self, dy, w, x_shape, stride, pad, addend, stable):
        return tb.conv_dgrad(dy, w, x_shape, stride, pad, addend)

    def _emu_conv_dgrad_bnbwd(self, dy, w, x_shape, stride, pad, addend, stable, bn_out, bn_yraw, mean, invstd, pre, cap6):
        assert not (cap6 and bn_out is None)
        relu = 2 if cap6 else (1 if bn_out is not None else 0)
        dx, sums = tb.conv_dgrad_bnbwd(dy, w, x_shape, stride, pad, addend, bn_out, bn_yraw, mean, invstd, relu)
        return dx, self._stats_out(sums, pre)

    def _emu_conv_wgrad(self, dy, x, out_grad, r, stride, pad, accumulate, prezeroed, k_valid, k_ld):
        if k_valid == 0:
            gw = torch.ops.aten.convolution_backward(dy, x, dy.new_empty(out_grad.shape), None, [stride, stride], [pad, pad],
                                                     [1, 1], False, [0, 0], 1, [False, True, False])[1].float()
        else:                                      # stem: dy [M, Cout, 1, 1], im2col matrix [M, Kp, 1, 1]
            cout, cin, R, S = out_grad.shape
            assert k_valid == cin * R * S == k_ld
            m = dy.shape[0]
            g2 = dy.reshape(m, cout).float().t() @ x.reshape(m, -1).float()[:, :k_valid]       # K order (r, s, cin)
            gw = g2.view(cout, R, S, cin).permute(0, 3, 1, 2)
        if accumulate or prezeroed:                # prezeroed: the kernel adds into a buffer it is told is zero
            out_grad.add_(gw)
        else:
            out_grad.copy_(gw)

    # ---- depthwise
    def _emu_dwconv_fwd(self, x, w, stride, want_stats, pre):
        y, sums = tb.dwconv_fwd(x, w, stride, want_stats)
        return y, (self._stats_out(sums, pre) if want_stats else None)

    def _emu_dwconv_dgrad(self, dy, w, x_shape, stride):
        return tb.dwconv_dgrad(dy, w, x_shape, stride)

    def _emu_dwconv_dgrad_bnbwd(self, dy, w, x_shape, stride, bn_out, bn_yraw, mean, invstd, pre, cap6):
        relu = 2 if cap6 else (1 if bn_out is not None else 0)
        dx, sums = tb.dwconv_dgrad_bnbwd(dy, w, x_shape, stride, bn_out, bn_yraw, mean, invstd, relu)
        return dx, self._stats_out(sums, pre)

    def _emu_dwconv_wgrad(self, dy, x, out_grad, stride, accumulate, prezeroed):
        tb.dwconv_wgrad(dy, x, stride, out_grad, accumulate or prezeroed)

    # ---- BatchNorm
    def _emu_channel_sums(self, y):
        yf = y.float()
        return torch.stack([yf.sum(dim=(0, 2, 3)), (yf * yf).sum(dim=(0, 2, 3))])

    def _emu_bn_act_fwd(self, y_raw, sums, gamma, beta, rmean, rvar, momentum, eps, residual, relu, training):
        assert relu in (0, 1, 2) and isinstance(relu, int)
        return tb.bn_act_fwd(y_raw, sums.view(2, -1) if training else None, gamma, beta, rmean, rvar, momentum, eps,
                             residual, relu, training)

    def _emu_bn_act_bwd(self, dout, out, y_raw, mean, invstd, gamma, relu, has_res, dg, db, ag, ab, scratch, sums_ready):
        c = y_raw.shape[1]
        if sums_ready:
            assert scratch is not None
        elif scratch is not None:
            assert float(scratch.abs().sum()) == 0.0, "BN-backward scratch handed over as zeroed is not zero"
        dy, _, _, dres = tb.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, relu, has_res, tb.GradSlot(dg, ag),
                                       tb.GradSlot(db, ab), sums=scratch.reshape(-1)[:2 * c].view(2, c) if sums_ready else None)
        return dy, dres

    def _emu_bn_act_bwd_res(self, dout, out, y_raw, mean, invstd, gamma, relu, dg, db, ag, ab, scratch, sums_ready,
                            res_yraw, res_mean, res_invstd, pre):
        c = y_raw.shape[1]
        dy, dres, rs = tb.bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma, relu, tb.GradSlot(dg, ag), tb.GradSlot(db, ab),
                                         scratch.reshape(-1)[:2 * c].view(2, c) if sums_ready else None, res_yraw,
                                         res_mean, res_invstd)
        return dy, dres, self._stats_out(rs, pre)

    # ---- pooling / head / optimizer / bookkeeping
    def _emu_maxpool_fwd(self, x, want_idx):
        y, aux = tb.maxpool_fwd(x, want_idx)
        return y, (aux[0] if want_idx else None)

    def _emu_maxpool_bwd(self, dy, idx, shape):
        return tb.maxpool_bwd(dy, (idx, tuple(shape)))

    def _emu_maxpool_bwd_bn(self, dy, idx, shape, bn_out, bn_yraw, mean, invstd, pre):
        dx, sums = tb.maxpool_bwd_bn(dy, (idx, tuple(shape)), bn_out, bn_yraw, mean, invstd, 1 if bn_out is not None else 0)
        return dx, self._stats_out(sums, pre)

    def _emu_head_fwd_bwd(self, feat, fc_w, fc_b, labels, loss_scale, n_valid, dw_out, db_out, accumulate, need_dfeat, scratch):
        return tb.head_fwd_bwd(feat, fc_w, fc_b, labels, loss_scale, n_valid, dw_out, db_out, accumulate, need_dfeat)

    def _emu_adam_step(self, master, grad, m, v, shadow, step_t, lr, b1, b2, eps, grad_scale, prev, diff, zero_grad,
                       live_blocks, bump, max_ctas):
        if live_blocks is None:
            tb.adam_step(master, grad, m, v, shadow, step_t, lr, b1, b2, eps, grad_scale, prev, zero_grad, live_blocks, diff,
                         bump, max_ctas)
            return
        # dead-block elision: the kernel only visits the listed 64-element blocks (gather, update, scatter)
        idx = _wire_index(live_blocks, master.numel())
        parts = [t[idx].contiguous() if t is not None else None for t in (master, grad, m, v, prev)]
        sh = shadow[idx].contiguous() if shadow is not None else None
        tb.adam_step(parts[0], parts[1], parts[2], parts[3], sh, step_t, lr, b1, b2, eps, grad_scale, parts[4], zero_grad, None,
                     diff, bump, max_ctas)
        for dst, src in zip((master, grad, m, v, prev), parts):
            if dst is not None:
                dst[idx] = src
        if shadow is not None:
            shadow[idx] = sh

    def _emu_grad_diff_sq(self, grad, prev):
        return tb.grad_diff_sq(grad, prev)

    def _emu_u8_normalize(self, images, mean, std):
        return tb.stem_prepare(images, mean, std, BF16)

    def _emu_stats_update(self, stats, has_prev, loss, correct, batch, diff_sq):
        tb.stats_update(stats, has_prev, loss, correct, batch, diff_sq)


def _wire_index(live, n):
    """element offsets, in wire order, of a bucket with dead-block compaction (`live`: int32 indices of 64-element blocks)"""
    if live is None:
        return torch.arange(n)
    return (live.long()[:, None] * 64 + torch.arange(64)[None, :]).reshape(-1)


class FakePeerComm:
    """Emulation of horizonml_b200._C.PeerComm for ranks that live in ONE process (the virtual-rank tests): the k-th
    collective of a group completes when the last rank has issued its k-th call (every rank queues its calls in order, like
    kernels on its stream) — the tests issue every rank's call before they look at any
    result, as they must on a GPU (the kernels of the early ranks spin until the late ones arrive).  Arithmetic follows
    the kernels' contract (csrc/comm.cu): wire = bf16(grad * scale) or fp32, summed in rank order in fp32; two-shot and
    NVLS round the sum to the wire type for their second hop; ZeRO-1: rank r owns wire vectors [r*q, (r+1)*q), q =
    ceil(nv / W)."""
    _real = None

    def __init__(self, rank, world, device, max_wire_bytes, max_blocks, heap_bytes=0):
        try:                                       # the real constructor must accept the argument list
            FakePeerComm._real(rank, world, device, max_wire_bytes, max_blocks, heap_bytes)
        except TypeError as e:
            raise AssertionError(f"C.PeerComm: argument list rejected by the binding: {e}") from None
        except RuntimeError:
            pass                                   # (no device to create the communicator on)
        self.rank, self.world, self.max_wire_bytes = rank, world, max_wire_bytes
        self.group, self.queue, self.cap = None, [], 0      # queue: this rank's issued, not yet completed collectives (stream order)

    @staticmethod
    def link_local(comms):
        assert sorted(c.rank for c in comms) == list(range(len(comms))) and all(c.world == len(comms) for c in comms)
        for c in comms:
            c.group = sorted(comms, key=lambda x: x.rank)

    def set_block_cap(self, cap):
        self.cap = int(cap)

    def error(self):
        return 0

    def zero1_shard(self, n_wire):
        return ((n_wire // 8 + self.world - 1) // self.world) * 8

    def blocks_for(self, n, algo, wire_bf16):
        return 1

    def _arrive(self, kind, payload):
        assert self.group is not None, "communicator not linked"
        self.queue.append((kind, payload))
        while all(c.queue for c in self.group):            # the k-th collective of every rank has been issued: run it
            heads = [c.queue.pop(0) for c in self.group]
            assert len({h[0] for h in heads}) == 1, "ranks issued different collectives: " + str([h[0] for h in heads])
            getattr(FakePeerComm, "_do_" + heads[0][0])(self.group, [h[1] for h in heads])

    def allreduce(self, grad, algo, wire_bf16, scale, live_blocks=None):
        assert algo in ("oneshot", "twoshot", "nvls", "ll", "bulk") and grad.dtype == torch.float32
        n = live_blocks.numel() * 64 if live_blocks is not None else grad.numel()
        assert n % (8 if wire_bf16 else 4) == 0 and n * (2 if wire_bf16 else 4) <= self.max_wire_bytes
        self._arrive("allreduce", (grad, algo, wire_bf16, scale, live_blocks))

    @staticmethod
    def _do_allreduce(group, calls):
        _, algo, wire_bf16, scale, live = calls[0]
        idx = _wire_index(live, calls[0][0].numel())
        total = torch.zeros(idx.numel())
        for g, *_ in calls:                        # rank order
            w = g[idx] * scale
            total += w.bfloat16().float() if wire_bf16 else w
        if algo in ("twoshot", "nvls") and wire_bf16:
            total = total.bfloat16().float()       # the reduced slice travels as bf16 on the second hop
        for g, *_ in calls:
            g[idx] = total

    def zero1_step(self, grad, scale, live_blocks, master, m, v, shadow, prev, diff_out, step, lr, b1, b2, eps, bump):
        n = live_blocks.numel() * 64 if live_blocks is not None else grad.numel()
        shard = self.zero1_shard(n)
        assert m.numel() == shard == v.numel() and (prev is None or prev.numel() == shard) and shadow.dtype == BF16
        self._arrive("zero1", (grad, scale, live_blocks, master, m, v, shadow, prev, diff_out, step, lr, b1, b2, eps, bump))

    @staticmethod
    def _do_zero1(group, calls):
        W = len(group)
        grad0, scale, live = calls[0][0], calls[0][1], calls[0][2]
        idx = _wire_index(live, grad0.numel())
        n = idx.numel()
        total = torch.zeros(n)
        for c in calls:
            total += (c[0][idx] * scale).bfloat16().float()
            c[0][idx] = 0.0                        # the pack pass clears the gradient
        q = ((n // 8 + W - 1) // W) * 8
        new_params = torch.empty(n)
        diffs = []
        for r, c in enumerate(calls):
            _, _, _, master, m, v, shadow, prev, diff_out, step, lr, b1, b2, eps, bump = c
            lo, hi = min(r * q, n), min(r * q + q, n)
            own, k = idx[lo:hi], hi - lo
            g = total[lo:hi]
            if prev is not None and diff_out is not None:
                diffs.append(((g - prev[:k]) ** 2).sum())
                prev[:k] = g
            t = float(step[0]) + 1.0
            m[:k] = b1 * m[:k] + (1 - b1) * g
            v[:k] = b2 * v[:k] + (1 - b2) * g * g
            denom = v[:k].sqrt() / (1 - b2 ** t) ** 0.5 + eps
            master[own] = master[own] - (lr / (1 - b1 ** t)) * m[:k] / denom
            new_params[lo:hi] = master[own]
        for c in calls:
            c[6][idx] = new_params.to(BF16)        # every rank's shadow gets every slice
            if c[7] is not None and c[8] is not None:
                c[8].add_(sum(diffs))
            if c[14]:
                c[9].add_(1.0)


@pytest.fixture
def shim(monkeypatch):
    real = _ext.load(required=False)
    if real is None:
        pytest.skip("extension not built")
    import horizonml_b200.ops.native_backend as nb
    from horizonml_b200.ops import functional as fn
    s = ShimC(real)
    monkeypatch.setattr(nb, "C", s)
    monkeypatch.setattr(nb, "_dev", lambda t: True)
    monkeypatch.setattr(nb, "_STRICT", True)                      # any fallback raises
    state = {"native": False}
    monkeypatch.setattr(fn, "_be", lambda t: nb if state["native"] else tb)
    return s, nb, state


def _train_steps(make_model, images, labels, nb, state, native, steps=2, arena=True):
    """`steps` optimizer steps of model.forward_loss through FlatParams / FlatAdam; returns (losses, first-step flat
    gradient, final master parameters)."""
    from horizonml_b200 import ops
    from horizonml_b200.models.flat import FlatAdam, FlatParams
    state["native"] = native
    torch.manual_seed(11)                       # MobileNetV2's classifier dropout draws the same mask in both runs
    dev = torch.device("cpu")
    model = make_model().train()
    with torch.no_grad():                       # wide BatchNorm outputs: plenty of activations beyond ReLU6's cap, so a
        for n, p in model.named_parameters():   # wrong activation code anywhere changes the gradients visibly
            if p.dim() == 1 and n.endswith("weight"):
                p.mul_(4.0)
    flat = FlatParams(list(model.named_parameters()), dev, BF16)
    opt = FlatAdam(flat, lr=1e-3)
    losses, g0 = [], None
    for it in range(steps):
        x = ops.stem_prepare(images.permute(0, 3, 1, 2), dtype=BF16)
        if native and arena:
            nb.step_begin(dev)
        flat.begin_step()
        loss, _ = model.forward_loss(x, labels)
        loss.backward()
        ops.join_side()
        if native and arena:
            nb.step_end()
        if it == 0:
            g0 = flat.grad.clone()
        opt.step()
        losses.append(float(loss.detach()))
    return losses, g0, flat.master.clone()


def _close(a, b, tol):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)) < tol


@pytest.mark.parametrize("handoff", [False, True])
def test_resnet18_step_through_the_native_backend_on_a_shim(shim, handoff):
    import horizonml_b200.models.resnet as R
    s, nb, state = shim
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (16, 32, 32, 3), dtype=torch.uint8, generator=g)
    labels = torch.randint(0, 10, (16,), generator=g)
    old = R._BN_BWD_IN_DGRAD
    try:
        R._BN_BWD_IN_DGRAD = False
        ref = _train_steps(lambda: R.resnet18(10, seed=0), images, labels, nb, state, native=False)
        R._BN_BWD_IN_DGRAD = handoff
        before = Counter(nb.LAUNCHES)
        got = _train_steps(lambda: R.resnet18(10, seed=0), images, labels, nb, state, native=True)
    finally:
        R._BN_BWD_IN_DGRAD = old
        state["native"] = False
    assert not nb.FALLBACKS or sum(nb.FALLBACKS.values()) == 0, dict(nb.FALLBACKS)
    assert abs(got[0][0] - ref[0][0]) < 2e-3 and abs(got[0][1] - ref[0][1]) < 5e-2, (got[0], ref[0])
    assert _close(got[1], ref[1], 2e-2), "first-step gradient differs from the PyTorch-op backend"
    assert _close(got[2], ref[2], 5e-2)           # (Adam's first steps are sign-like: near-zero gradients may flip)
    per_step = {k: (nb.LAUNCHES[k] - before[k]) // 2 for k in nb.LAUNCHES}
    # 20 convs (the stem through im2col + the 1x1 GEMM), 20 BatchNorms, pool, head, optimizer: every op went through a binding
    assert s.calls["conv_fwd"] == 2 * 20 and s.calls["stem_pack"] == 2 and s.calls["conv_wgrad"] == 2 * 20
    assert s.calls["im2col_small"] == 0                      # the forward's im2col matrix is reused by the stem's wgrad
    assert s.calls["bn_act_fwd"] == 2 * 20 and s.calls["head_fwd_bwd"] == 2 and s.calls["adam_step"] >= 2
    assert s.calls["u8_normalize"] == 2 and s.calls["maxpool_fwd"] == 2
    if handoff:
        assert s.calls["conv_dgrad_bnbwd"] == 2 * 15 and s.calls["maxpool_bwd_bn"] == 2 and s.calls["bn_act_bwd_res"] == 2 * 3
        assert s.calls["conv_dgrad"] == 2 * (19 - 15) and s.calls["bn_act_bwd"] == 2 * (20 - 3) and s.calls["maxpool_bwd"] == 0
        assert per_step["bn_act_bwd"] == 40 - 19
    else:
        assert s.calls["conv_dgrad"] == 2 * 19 and s.calls["bn_act_bwd"] == 2 * 20 and per_step["bn_act_bwd"] == 40
        assert s.calls["conv_dgrad_bnbwd"] == 0 and s.calls["bn_act_bwd_res"] == 0 and s.calls["maxpool_bwd"] == 2


@pytest.mark.parametrize("handoff", [False, True])
def test_mobilenet_step_through_the_native_backend_on_a_shim(shim, handoff):
    import horizonml_b200.models.resnet as R
    from horizonml_b200.models.mobilenet import mobilenet_v2
    s, nb, state = shim
    g = torch.Generator().manual_seed(1)
    images = torch.randint(0, 256, (8, 32, 32, 3), dtype=torch.uint8, generator=g)
    labels = torch.randint(0, 10, (8,), generator=g)
    old = R._BN_BWD_IN_DGRAD
    try:
        R._BN_BWD_IN_DGRAD = False
        ref = _train_steps(lambda: mobilenet_v2(10, seed=0), images, labels, nb, state, native=False, steps=1)
        R._BN_BWD_IN_DGRAD = handoff
        got = _train_steps(lambda: mobilenet_v2(10, seed=0), images, labels, nb, state, native=True, steps=1)
    finally:
        R._BN_BWD_IN_DGRAD = old
        state["native"] = False
    assert sum(nb.FALLBACKS.values()) == 0, dict(nb.FALLBACKS)
    assert abs(got[0][0] - ref[0][0]) < 5e-3, (got[0], ref[0])
    assert _close(got[1], ref[1], 3e-2)
    assert s.calls["dwconv_fwd"] == 17 and s.calls["dwconv_wgrad"] == 17 and s.calls["bn_act_fwd"] == 52
    if handoff:
        assert s.calls["dwconv_dgrad_bnbwd"] == 16 and s.calls["conv_dgrad_bnbwd"] == 34
    else:
        assert s.calls["dwconv_dgrad"] == 17 and s.calls["dwconv_dgrad_bnbwd"] == 0 and s.calls["conv_dgrad_bnbwd"] == 0


@pytest.mark.parametrize("model_name", ["resnet18", "mobilenet"])
def test_dp_engine_steps_through_the_native_backend_on_a_shim(shim, monkeypatch, model_name):
    """The data-parallel engine itself (trainers/dp.py: stem preparation, statistics arena per step, flat-bucket
    gradients, fused Adam with the gradient-divergence bookkeeping, on-device step statistics) on the shimmed native
    backend: same loss curve and parameters as the engine on the PyTorch-op backend."""
    from horizonml_b200.config import TrainConfig
    from horizonml_b200.ops import functional as fn
    from horizonml_b200.trainers.common import Runtime
    from horizonml_b200.trainers.dp import DPEngine
    s, nb, state = shim
    dev = torch.device("cpu")
    monkeypatch.setattr(fn, "step_begin", lambda device=None: nb.step_begin(dev) if state["native"] else None)
    monkeypatch.setattr(fn, "step_end", lambda: nb.step_end() if state["native"] else None)
    g = torch.Generator().manual_seed(3)
    xs = torch.randint(0, 256, (16, 32, 32, 3), dtype=torch.uint8, generator=g)
    ys = torch.randint(0, 10, (16,), generator=g)
    res = {}
    for native in (False, True):
        state["native"] = native
        be = "native" if native else "torch"
        cfg = TrainConfig(strategy="data", world_size=1, batch_size=16, device="cpu", dtype="bf16", backend=be,
                          model=model_name, quiet=True, cuda_graph=False)
        torch.manual_seed(9)                                 # (MobileNetV2's dropout mask)
        eng = DPEngine(cfg, Runtime(0, 1, dev, BF16, be, "none"))
        for _ in range(3):
            eng.step(xs, ys)
        res[native] = (eng.stats.buf.clone(), eng.flat.master.clone())
    state["native"] = False
    assert sum(nb.FALLBACKS.values()) == 0, dict(nb.FALLBACKS)
    assert s.calls["stats_update"] == 3 and s.calls["u8_normalize"] == 3 and s.calls["adam_step"] >= 3
    (st0, m0), (st1, m1) = res[False], res[True]
    assert abs(float(st0[0] - st1[0])) < 0.1 * abs(float(st0[0])) and float(st0[2]) == float(st1[2]) == 48.0   # Σ loss, samples
    assert float(st1[4]) == 2.0 and float(st1[3]) > 0           # gradient divergence accumulated over steps 2 and 3
    assert abs(float(st1[3] - st0[3])) < 0.2 * float(st0[3])
    assert _close(m1, m0, 5e-2)


@pytest.mark.parametrize("model_name", ["resnet18", "mobilenet"])
def test_eval_forward_through_the_native_backend_on_a_shim(shim, model_name):
    """Inference path (running statistics, no autograd): conv / depthwise conv / BN-apply / pool bindings called with
    training=False, logits equal to the PyTorch-op backend's."""
    import horizonml_b200.models.resnet as R
    from horizonml_b200 import ops
    from horizonml_b200.models.flat import FlatParams
    from horizonml_b200.models.mobilenet import mobilenet_v2
    s, nb, state = shim
    g = torch.Generator().manual_seed(5)
    images = torch.randint(0, 256, (8, 32, 32, 3), dtype=torch.uint8, generator=g)
    out = {}
    for native in (False, True):
        state["native"] = native
        model = (R.resnet18(10, seed=0) if model_name == "resnet18" else mobilenet_v2(10, seed=0)).eval()
        with torch.no_grad():
            for n, b in model.named_buffers():            # non-trivial running statistics
                if n.endswith("running_mean"):
                    b.copy_(torch.linspace(-0.2, 0.2, b.numel()))
                elif n.endswith("running_var"):
                    b.copy_(torch.linspace(0.5, 1.5, b.numel()))
        FlatParams(list(model.named_parameters()), torch.device("cpu"), BF16)
        with torch.no_grad():
            x = ops.stem_prepare(images.permute(0, 3, 1, 2), dtype=BF16)
            out[native] = model(x).float()
    state["native"] = False
    assert sum(nb.FALLBACKS.values()) == 0, dict(nb.FALLBACKS)
    assert s.calls["bn_act_fwd"] == (20 if model_name == "resnet18" else 52) and s.calls["channel_sums"] == 0
    assert out[True].shape == (8, 10) and _close(out[True], out[False], 2e-2), (out[True] - out[False]).abs().max()


@pytest.mark.deep
@pytest.mark.parametrize("which,world,handoff", [("dp", 2, False), ("dpz", 2, True), ("pp", 2, True), ("tp", 2, False),
                                                 ("tp", 2, True)])
def test_parallel_engines_on_a_shim(which, world, handoff):
    """trainers/dp.py at two ranks (bucketed gradient all-reduce over gloo; `dpz`: ZeRO-1 in its torch.distributed form),
    trainers/pp.py (1F1B runner, 4 micro-batches, activations / gradients over gloo p2p at two stages) and trainers/tp.py
    (channel-split layer3/4 blocks and the tensor-parallel head on their non-fused paths: 128- / 256-channel shards at
    world 2) with every op on the shimmed native backend, one process per rank: same step statistics (loss, correct,
    samples, gradient divergence) as the same engine on the PyTorch-op backend, no fallback — with and without the
    BatchNorm hand-offs (which PP shares with the dense model and TP uses inside its blocks)."""
    import json
    import subprocess
    import sys
    if _ext.load(required=False) is None:
        pytest.skip("extension not built")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, os.path.join(here, "helpers_shim_engines.py"), which, str(world)] + (["handoff"] if handoff else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("SHIM_ENGINE ")]
    assert r.returncode == 0 and len(lines) == world, (r.stdout[-1500:], r.stderr[-3000:])
    ranks = [json.loads(ln[len("SHIM_ENGINE "):]) for ln in lines]
    passes = 8 if which == "pp" else 2                    # forward/backward passes per rank (micro-batches x steps)
    total = {}
    for res in ranks:
        assert sum(res["fallbacks"].values()) == 0, res["fallbacks"]
        for k, v in res["calls"].items():
            total[k] = total.get(k, 0) + v
    last = ranks[-1]                                      # the rank that owns the loss (last stage; every TP rank)
    t, n = last["torch"], last["native"]
    assert abs(t[0] - n[0]) < 0.02 * abs(t[0]) and t[1:3] == n[1:3] and t[4:] == n[4:], (t, n)     # loss; correct, samples; counts
    assert abs(t[3] - n[3]) < 0.1 * max(t[3], 1e-6), (t, n)                                       # gradient divergence
    layers = 20 * passes * (1 if which == "pp" else world)      # PP: the 20 convs are spread over the stages
    assert total["conv_fwd"] == layers and total["conv_wgrad"] == layers and total["bn_act_fwd"] == layers, total
    if handoff:
        assert total.get("conv_dgrad_bnbwd", 0) >= 8 * passes, total
    else:
        assert total.get("conv_dgrad_bnbwd", 0) == 0 and total.get("bn_act_bwd_res", 0) == 0, total


def test_smoke_sequence_without_the_statistics_arena_on_a_shim(shim):
    """__graft_entry__.smoke() steps the model without ops.step_begin(): no pre-zeroed statistics arena, every binding gets
    None for its scratch / pre-zeroed buffers and has to allocate and clear its own — same gradients as with the arena."""
    import horizonml_b200.models.resnet as R
    s, nb, state = shim
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (16, 32, 32, 3), dtype=torch.uint8, generator=g)
    labels = torch.randint(0, 10, (16,), generator=g)
    try:
        ref = _train_steps(lambda: R.resnet18(10, seed=0), images, labels, nb, state, native=False)
        got = _train_steps(lambda: R.resnet18(10, seed=0), images, labels, nb, state, native=True, arena=False)
    finally:
        state["native"] = False
    assert sum(nb.FALLBACKS.values()) == 0 and not nb.ARENA.active
    assert abs(got[0][0] - ref[0][0]) < 2e-3 and _close(got[1], ref[1], 2e-2) and _close(got[2], ref[2], 5e-2)
