"""Native op backend: every function launches hand-written sm_100a kernels from ``csrc/``.

Same signatures as ``torch_backend``.  Convolutions run on the tcgen05/TMEM/TMA implicit-GEMM
kernels; RGB stems (ResNet 7×7/2, MobileNetV2 3×3/2) go im2col → the same GEMM kernel (TMA needs ≥16-byte rows);
depthwise 3×3 convolutions run on the SIMT kernels of ``csrc/depthwise.cu``; shapes the tiles cannot express
(channels not a multiple of 8, odd spatial sizes under stride 2) are routed to the PyTorch oracle and counted in
``FALLBACKS`` — ``HZ_STRICT_NATIVE=1`` turns that into an error.
"""
from __future__ import annotations

import os
from collections import Counter
from typing import Optional

import torch

from . import _ext
from . import torch_backend as _tb

C = _ext.load(required=True)
FALLBACKS: Counter = Counter()
LAUNCHES: Counter = Counter()      # native kernel launches issued from Python (bench.py reports these)
_STRICT = os.environ.get("HZ_STRICT_NATIVE", "0") == "1"
STEM_KP = 192                       # 7*7*3 = 147 padded to 3 k-blocks of 64


class _Arena:
    """Pre-zeroed fp32 scratch for per-channel statistics (conv-epilogue BN sums, BN-backward sums):
    one fill per step instead of one cudaMemset node per layer."""

    def __init__(self, n=1 << 18):
        self.n, self.buf, self.off, self.used_prev, self.active = n, None, 0, 0, False

    def begin(self, device):
        if self.buf is None or self.buf.device != device:
            self.buf = torch.zeros(self.n, dtype=torch.float32, device=device)
            self.high = 0
        self.high = max(getattr(self, "high", 0), self.off)
        if self.high > 0:
            self.buf[:self.high].zero_()          # everything ever handed out (slices baked into graphs too)
        self.used_prev, self.off, self.active = self.off, 0, True

    def take(self, rows, c, device):
        if not self.active or self.buf is None or self.buf.device != device:
            return None
        n = rows * c
        if self.off + n > self.n:
            return None
        t = self.buf[self.off:self.off + n].view(rows, c)
        self.off += (n + 31) // 32 * 32
        return t


ARENA = _Arena()
_STEM_CACHE = {"key": None, "A": None}


def step_begin(device=None):
    ARENA.begin(torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()))
    _STEM_CACHE["key"] = None
    _STEM_CACHE["A"] = None


def step_end():
    ARENA.active = False


def _fallback(name: str, why: str):
    FALLBACKS[name] += 1
    if _STRICT:
        raise RuntimeError(f"native backend fallback in {name}: {why}")


def _dev(t: torch.Tensor) -> bool:
    """Tensor lives where the kernels run (one seam: tests/test_cpu_native_plumbing.py drives this module on CPU
    tensors against a shim of the extension)."""
    return t.is_cuda


def _bf16_cl(t: torch.Tensor) -> bool:
    return _dev(t) and t.dtype == torch.bfloat16 and t.dim() == 4


def _conv_ok(x_shape, w_shape, stride, pad) -> bool:
    n, cin, h, w = x_shape
    cout, _, r, s = w_shape
    return r == s and bool(C.conv_supported(n, h, w, cin, cout, r, stride, pad))


def _is_stem(x_shape, w_shape) -> bool:
    """A dense conv over fewer than 8 input channels (RGB stems: ResNet 7x7/2 -> 64, MobileNetV2 3x3/2 -> 32): im2col
    to [M, Kp] + the 1x1 tcgen05 GEMM."""
    return x_shape[1] < 8 and w_shape[0] % 8 == 0 and w_shape[1] * w_shape[2] * w_shape[3] <= STEM_KP


def _stem_kp(w_shape) -> int:
    """Padded K of the im2col matrix: 192 for the 7x7x3 = 147 stem (the one-launch stem_pack kernel), else the next
    multiple of 64 (one k-block of the GEMM per 64)."""
    k = w_shape[1] * w_shape[2] * w_shape[3]
    return STEM_KP if k == 147 else (k + 63) // 64 * 64


# ------------------------------------------------------------------------------------------------
def _stable(w) -> bool:
    """True for the persistent bf16 shadow of a FlatParams parameter: last written by the optimizer, i.e. at least
    two kernels before any conv that reads it — the kernel may then request weight tiles before
    ``griddepcontrol.wait`` (programmatic dependent launch).  Anything else (casts, slices, padded copies) may have
    been produced by the immediately preceding kernel and is only read after the wait."""
    if not getattr(w, "_hz_stable", False):
        return False
    # Early launch is transitive (every kernel issues launch_dependents at its top), so in EAGER mode — warm-up steps,
    # --no_cuda_graph, an off-size last batch — a conv of the next step could prefetch shadow weights while this step's
    # Adam is still writing them.  Inside a captured graph the optimizer is always >= 2 kernels upstream of the first
    # conv that reads a weight, within and across replays, so the prefetch is only enabled there.
    return bool(w.is_cuda and torch.cuda.is_current_stream_capturing())


def conv_fwd(x, w, stride: int, pad: int, want_stats: bool):
    if _bf16_cl(x) and w.dtype == torch.bfloat16:
        if _conv_ok(x.shape, w.shape, stride, pad):
            LAUNCHES["conv_fwd"] += 1
            pre = ARENA.take(2, w.shape[0], x.device) if want_stats else None
            y, stats = C.conv_fwd(x, w, stride, pad, want_stats, pre, _stable(w))
            return y, (stats if want_stats else None)
        if _is_stem(x.shape, w.shape):
            n, cin, h, wd = x.shape
            cout, _, r, _ = w.shape
            ho, wo = (h + 2 * pad - r) // stride + 1, (wd + 2 * pad - r) // stride + 1
            w2d = w.permute(0, 2, 3, 1).reshape(cout, -1)
            if not w2d.is_contiguous():
                w2d = w2d.contiguous()
            kp = _stem_kp(w.shape)
            A, wp = C.stem_pack(x, w2d, r, stride, pad, kp)       # [N*Ho*Wo, Kp], [Cout, Kp] (7x7x3: one launch)
            _STEM_CACHE["key"], _STEM_CACHE["A"] = (x.data_ptr(), x._version, tuple(x.shape)), A
            LAUNCHES["stem_im2col"] += 1 if kp == STEM_KP else 2
            LAUNCHES["conv_fwd"] += 1
            pre = ARENA.take(2, cout, x.device) if want_stats else None
            y2, stats = C.conv_fwd(A.view(-1, kp, 1, 1), wp.view(cout, kp, 1, 1), 1, 0, want_stats, pre,
                                  False)          # wp was produced by the kernel right before: no early weight prefetch
            y = y2.reshape(n, ho, wo, cout).permute(0, 3, 1, 2)
            return y, (stats if want_stats else None)
    _fallback("conv_fwd", f"x={tuple(x.shape)} w={tuple(w.shape)} s={stride}")
    return _tb.conv_fwd(x, w, stride, pad, want_stats)


# Measured on B200 (profiles/README.md): the device-wide barrier costs more than the separate BN kernel saves once
# programmatic dependent launch overlaps that kernel's launch and prologue (0.586 vs 0.571 ms/step) — opt-in.
_FUSE_BN = os.environ.get("HZ_FUSE_BN", "0") == "1"
# Same finding for BN backward: reduce → device-wide barrier → apply in one kernel (0.572 ms/step) loses to two
# PDL-chained kernels (0.562 ms/step); HZ_BN_BWD_FUSED=1 selects the single-kernel variant.
_BN_BWD_FUSED = os.environ.get("HZ_BN_BWD_FUSED", "0") == "1"


def conv_bn_act_fwd(x, w, stride: int, pad: int, gamma, beta, rmean, rvar, momentum, eps, residual, relu: bool):
    """Training-mode conv → BN(batch stats) → (+residual) → (ReLU) as ONE kernel: the conv epilogue leaves Σy, Σy²
    in the statistics arena, a device-wide barrier makes them final, and every CTA normalises the tile it still
    holds in shared memory.  Returns (y_raw, out, mean, invstd), or None when not applicable (caller runs the
    conv and BN kernels separately)."""
    if not (_FUSE_BN and int(relu) < 2 and _bf16_cl(x) and w.dtype == torch.bfloat16 and C.channel_ok(w.shape[0]) == 1):
        return None
    if residual is not None and not _bf16_cl(residual):
        return None
    cout = w.shape[0]
    direct = _conv_ok(x.shape, w.shape, stride, pad)
    if not direct and not _is_stem(x.shape, w.shape):
        return None
    scratch = ARENA.take(1, 2 * cout + 32, x.device)          # [Σy | Σy² | barrier counter], pre-zeroed
    if scratch is None:
        return None
    LAUNCHES["conv_bn_fused"] += 1
    if direct:
        y, out, mean, invstd = C.conv_bn_act_fwd(x, w, stride, pad, scratch.view(-1), gamma, beta, rmean, rvar,
                                                 momentum, eps, residual, relu, _stable(w))
        return y, out, mean, invstd
    n, cin, h, wd = x.shape
    r = w.shape[2]
    ho, wo = (h + 2 * pad - r) // stride + 1, (wd + 2 * pad - r) // stride + 1
    w2d = w.permute(0, 2, 3, 1).reshape(cout, -1)
    if not w2d.is_contiguous():
        w2d = w2d.contiguous()
    kp = _stem_kp(w.shape)
    A, wp = C.stem_pack(x, w2d, r, stride, pad, kp)
    _STEM_CACHE["key"], _STEM_CACHE["A"] = (x.data_ptr(), x._version, tuple(x.shape)), A
    LAUNCHES["stem_im2col"] += 1
    res2 = residual.permute(0, 2, 3, 1).reshape(-1, cout, 1, 1) if residual is not None else None
    y2, o2, mean, invstd = C.conv_bn_act_fwd(A.view(-1, kp, 1, 1), wp.view(cout, kp, 1, 1), 1, 0,
                                             scratch.view(-1), gamma, beta, rmean, rvar, momentum, eps, res2, relu,
                                             False)
    return (y2.reshape(n, ho, wo, cout).permute(0, 3, 1, 2), o2.reshape(n, ho, wo, cout).permute(0, 3, 1, 2),
            mean, invstd)


def bn_act_fwd(y_raw, sums, gamma, beta, rmean, rvar, momentum, eps, residual, relu, training):
    if _bf16_cl(y_raw) and C.channel_ok(y_raw.shape[1]):
        if training and sums is None:
            sums = C.channel_sums(y_raw)
            LAUNCHES["channel_sums"] += 1
        if sums is None:
            sums = torch.empty(2, y_raw.shape[1], dtype=torch.float32, device=y_raw.device)
        LAUNCHES["bn_act_fwd"] += 1
        out, mean, invstd = C.bn_act_fwd(y_raw, sums, gamma, beta, rmean, rvar, momentum, eps, residual,
                                         int(relu), training)        # 0 none | 1 ReLU | 2 ReLU6
        return out, mean, invstd
    _fallback("bn_act_fwd", f"{tuple(y_raw.shape)} {y_raw.dtype}")
    return _tb.bn_act_fwd(y_raw, sums, gamma, beta, rmean, rvar, momentum, eps, residual, relu, training)


def bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, relu, has_residual, dgamma_slot=None, dbeta_slot=None, sums=None):
    """``sums``: the final [2, C] (Σg, Σg·x̂) left by ``conv_dgrad_bnbwd`` in the epilogue of the dgrad kernel that
    produced ``dout`` — the reduce pass is skipped, only the apply kernel runs."""
    if sums is not None and _bf16_cl(y_raw) and C.channel_ok(y_raw.shape[1]):
        c = y_raw.shape[1]
        if dgamma_slot is None:
            dg = torch.empty(c, dtype=torch.float32, device=y_raw.device)
            db = torch.empty(c, dtype=torch.float32, device=y_raw.device)
            ag = ab = False
        else:
            dg, ag, db, ab = dgamma_slot.t, dgamma_slot.acc, dbeta_slot.t, dbeta_slot.acc
        LAUNCHES["bn_act_bwd"] += 1
        dy, dres = C.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, int(relu), has_residual, dg, db, ag, ab, sums, True)
        return dy, dg, db, (dres if has_residual else None)
    if _bf16_cl(y_raw) and C.channel_ok(y_raw.shape[1]):
        c = y_raw.shape[1]
        if dgamma_slot is None:
            dg = torch.empty(c, dtype=torch.float32, device=y_raw.device)
            db = torch.empty(c, dtype=torch.float32, device=y_raw.device)
            ag = ab = False
        else:
            dg, ag, db, ab = dgamma_slot.t, dgamma_slot.acc, dbeta_slot.t, dbeta_slot.acc
        LAUNCHES["bn_act_bwd"] += 2
        # [Σg | Σg·x̂ | barrier counter], pre-zeroed: one kernel (reduce → device-wide barrier → apply); without
        # room for the counter the binding runs the reduce and apply kernels separately
        scratch = ARENA.take(1, 2 * c + (32 if _BN_BWD_FUSED else 0), y_raw.device)
        LAUNCHES["bn_act_bwd"] -= 1 if (scratch is not None and _BN_BWD_FUSED) else 0
        dy, dres = C.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, int(relu), has_residual, dg, db, ag, ab,
                                scratch, False)
        return dy, dg, db, (dres if has_residual else None)
    _fallback("bn_act_bwd", f"{tuple(y_raw.shape)}")
    return _tb.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, relu, has_residual, dgamma_slot, dbeta_slot)


def bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma, relu, dgamma_slot, dbeta_slot, sums, res_yraw, res_mean,
                   res_invstd):
    """bn_act_bwd of a layer whose residual comes from a BatchNorm without activation (a BasicBlock's downsample
    branch): the apply kernel also takes that BatchNorm's backward sums from the residual gradient it stores.
    ``sums``: this layer's own final sums if its consumer's dgrad already took them.  Returns (dy, dres, res_sums) or
    None when the shape / activation is not covered (caller runs the plain bn_act_bwd)."""
    c = y_raw.shape[1]
    if not (_bf16_cl(y_raw) and _bf16_cl(res_yraw) and tuple(res_yraw.shape) == tuple(y_raw.shape)
            and C.channel_ok(c) == 1 and int(relu) in (0, 1) and dgamma_slot is not None
            and res_mean.dtype == torch.float32 and res_invstd.dtype == torch.float32):
        return None
    dg, ag, db, ab = dgamma_slot.t, dgamma_slot.acc, dbeta_slot.t, dbeta_slot.acc
    pre = ARENA.take(2, c, y_raw.device)
    if sums is not None:
        LAUNCHES["bn_act_bwd"] += 1
        dy, dres, rs = C.bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma, int(relu), dg, db, ag, ab, sums, True,
                                        res_yraw, res_mean, res_invstd, pre)
    else:
        LAUNCHES["bn_act_bwd"] += 2
        scratch = ARENA.take(1, 2 * c, y_raw.device)
        dy, dres, rs = C.bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma, int(relu), dg, db, ag, ab, scratch, False,
                                        res_yraw, res_mean, res_invstd, pre)
    return dy, dres, rs


def conv_dgrad(dy, w, x_shape, stride: int, pad: int, addend=None):
    """``addend``: bf16 channels_last tensor of x's shape added in the kernel epilogue (the residual-branch
    gradient), replacing autograd's separate accumulation kernel."""
    if _bf16_cl(dy) and w.dtype == torch.bfloat16 and _conv_ok(tuple(x_shape), w.shape, stride, pad):
        LAUNCHES["conv_dgrad"] += 1
        if addend is not None and not (_bf16_cl(addend) and tuple(addend.shape) == tuple(x_shape)):
            return C.conv_dgrad(dy, w, list(x_shape), stride, pad, None, _stable(w)).add_(addend)
        return C.conv_dgrad(dy, w, list(x_shape), stride, pad, addend, _stable(w))
    _fallback("conv_dgrad", f"x={tuple(x_shape)} w={tuple(w.shape)}")
    return _tb.conv_dgrad(dy, w, x_shape, stride, pad, addend)


def conv_dgrad_bnbwd(dy, w, x_shape, stride: int, pad: int, addend, bn_out, bn_yraw, bn_mean, bn_invstd, relu):
    """conv_dgrad whose epilogue also takes the BatchNorm-backward sums of the layer that produced the conv's input
    (``bn_out`` / ``bn_yraw`` / statistics of that layer; ``relu`` its activation code: 0 none, 1 ReLU, 2 ReLU6).
    Returns (dx, sums[2, Cin]) or None when the fused form does not apply (caller runs the plain dgrad)."""
    if not (_bf16_cl(dy) and w.dtype == torch.bfloat16 and _conv_ok(tuple(x_shape), w.shape, stride, pad)
            and int(relu) in (0, 1, 2) and _bf16_cl(bn_yraw) and tuple(bn_yraw.shape) == tuple(x_shape)
            and C.channel_ok(x_shape[1]) and bn_mean.dtype == torch.float32 and bn_invstd.dtype == torch.float32):
        return None
    if addend is not None and not (_bf16_cl(addend) and tuple(addend.shape) == tuple(x_shape)):
        return None
    LAUNCHES["conv_dgrad"] += 1
    pre = ARENA.take(2, x_shape[1], dy.device)
    dx, sums = C.conv_dgrad_bnbwd(dy, w, list(x_shape), stride, pad, addend, _stable(w), bn_out if int(relu) else None,
                                  bn_yraw, bn_mean, bn_invstd, pre, int(relu) == 2)
    return dx, sums


def _flat_dw_ok(out_grad: torch.Tensor, w_shape) -> bool:
    """out_grad must be the [Cout,R,S,Cin]-physical fp32 view (what FlatParams hands out)."""
    cout, cin, r, s = w_shape
    return (out_grad.dtype == torch.float32 and _dev(out_grad) and
            out_grad.permute(0, 2, 3, 1).is_contiguous())


def conv_wgrad(dy, x, w_shape, stride: int, pad: int, out_grad: torch.Tensor, accumulate: bool,
               prezeroed: bool = False):
    cout, cin, r, s = w_shape
    if _bf16_cl(dy) and _bf16_cl(x) and _flat_dw_ok(out_grad, w_shape):
        if _conv_ok(x.shape, w_shape, stride, pad):
            LAUNCHES["conv_wgrad"] += 1
            C.conv_wgrad(dy, x, out_grad, r, stride, pad, accumulate, prezeroed, 0, 0)
            return
        if _is_stem(x.shape, w_shape):
            kp = _stem_kp(w_shape)
            if _STEM_CACHE["key"] == (x.data_ptr(), x._version, tuple(x.shape)) and _STEM_CACHE["A"].shape[1] == kp:
                A = _STEM_CACHE["A"]            # the forward's im2col matrix is still alive
            else:
                A = C.im2col_small(x, r, stride, pad, kp)
                LAUNCHES["stem_im2col"] += 1
            n, _, ho, wo = dy.shape
            dy2 = dy.permute(0, 2, 3, 1).reshape(-1, cout, 1, 1)     # [M, Cout,1,1] (NHWC rows)
            LAUNCHES["conv_wgrad"] += 1
            C.conv_wgrad(dy2, A.view(-1, kp, 1, 1), out_grad, 1, 1, 0, accumulate, prezeroed, cin * r * s, cin * r * s)
            return
    _fallback("conv_wgrad", f"x={tuple(x.shape)} w={tuple(w_shape)}")
    _tb.conv_wgrad(dy, x, w_shape, stride, pad, out_grad, accumulate)


# ---- depthwise 3x3 convolution, pad 1 (csrc/depthwise.cu): weight [C, 1, 3, 3]
def _dw_ok(x_shape, stride) -> bool:
    n, c, h, w = x_shape
    return bool(C.dwconv_ok(n, h, w, c, stride))


def dwconv_fwd(x, w, stride: int, want_stats: bool):
    if _bf16_cl(x) and w.dtype == torch.bfloat16 and _dw_ok(x.shape, stride):
        LAUNCHES["dwconv_fwd"] += 1
        pre = ARENA.take(2, x.shape[1], x.device) if want_stats else None
        y, stats = C.dwconv_fwd(x, w, stride, want_stats, pre)
        return y, (stats if want_stats else None)
    _fallback("dwconv_fwd", f"x={tuple(x.shape)} {x.dtype} s={stride}")
    return _tb.dwconv_fwd(x, w, stride, want_stats)


def dwconv_dgrad(dy, w, x_shape, stride: int):
    if _bf16_cl(dy) and w.dtype == torch.bfloat16 and _dw_ok(tuple(x_shape), stride):
        LAUNCHES["dwconv_dgrad"] += 1
        return C.dwconv_dgrad(dy, w, list(x_shape), stride)
    _fallback("dwconv_dgrad", f"x={tuple(x_shape)}")
    return _tb.dwconv_dgrad(dy, w, x_shape, stride)


def dwconv_dgrad_bnbwd(dy, w, x_shape, stride: int, bn_out, bn_yraw, bn_mean, bn_invstd, relu):
    """Depthwise dgrad whose kernel also takes the BatchNorm-backward sums of the layer that produced the conv's input
    (activation code 0 / 1 / 2).  Returns (dx, sums[2, C]) or None when the fused form does not apply."""
    if not (_bf16_cl(dy) and w.dtype == torch.bfloat16 and _dw_ok(tuple(x_shape), stride) and int(relu) in (0, 1, 2)
            and _bf16_cl(bn_yraw) and tuple(bn_yraw.shape) == tuple(x_shape) and bn_mean.dtype == torch.float32
            and bn_invstd.dtype == torch.float32):
        return None
    LAUNCHES["dwconv_dgrad"] += 1
    pre = ARENA.take(2, x_shape[1], dy.device)
    dx, sums = C.dwconv_dgrad_bnbwd(dy, w, list(x_shape), stride, bn_out if int(relu) else None, bn_yraw, bn_mean,
                                    bn_invstd, pre, int(relu) == 2)
    return dx, sums


def dwconv_wgrad(dy, x, stride: int, out_grad: torch.Tensor, accumulate: bool, prezeroed: bool = False):
    if (_bf16_cl(dy) and _bf16_cl(x) and _dw_ok(x.shape, stride) and _dev(out_grad) and out_grad.dtype == torch.float32
            and out_grad.dim() == 4 and out_grad.stride(0) == 9 and out_grad.stride(2) == 3 and out_grad.stride(3) == 1):
        LAUNCHES["dwconv_wgrad"] += 1
        C.dwconv_wgrad(dy, x, out_grad, stride, accumulate, prezeroed)
        return
    _fallback("dwconv_wgrad", f"x={tuple(x.shape)}")
    _tb.dwconv_wgrad(dy, x, stride, out_grad, accumulate, prezeroed)


def maxpool_fwd(x, want_aux: bool = False):
    if _bf16_cl(x) and x.shape[1] % 8 == 0:
        LAUNCHES["maxpool"] += 1
        y, idx = C.maxpool_fwd(x, want_aux)
        return y, (("native", idx, tuple(x.shape)) if want_aux else None)
    _fallback("maxpool_fwd", str(tuple(x.shape)))
    return _tb.maxpool_fwd(x, want_aux)


def maxpool_bwd(dy, aux):
    if isinstance(aux, tuple) and aux and isinstance(aux[0], str) and aux[0] == "native":
        LAUNCHES["maxpool"] += 1
        return C.maxpool_bwd(dy, aux[1], list(aux[2]))
    return _tb.maxpool_bwd(dy, aux)


def maxpool_bwd_bn(dy, aux, bn_out, bn_yraw, bn_mean, bn_invstd, relu):
    """max-pool backward whose kernel also takes the BatchNorm-backward sums of the layer that feeds the pool (its output
    has no other consumer).  Returns (dx, sums[2, C]) or None when the fused form does not apply."""
    if not (isinstance(aux, tuple) and aux and isinstance(aux[0], str) and aux[0] == "native" and _bf16_cl(dy)
            and int(relu) in (0, 1) and _bf16_cl(bn_yraw) and tuple(bn_yraw.shape) == tuple(aux[2])
            and C.channel_ok(aux[2][1]) == 1 and bn_mean.dtype == torch.float32 and bn_invstd.dtype == torch.float32):
        return None
    LAUNCHES["maxpool"] += 1
    pre = ARENA.take(2, aux[2][1], dy.device)
    dx, sums = C.maxpool_bwd_bn(dy, aux[1], list(aux[2]), bn_out if int(relu) else None, bn_yraw, bn_mean, bn_invstd, pre)
    return dx, sums


def head_fwd_bwd(feat, fc_w, fc_b, labels, loss_scale, n_valid, dw_out, db_out, accumulate, need_dfeat=True):
    if (_bf16_cl(feat) and fc_w.dtype == torch.float32 and fc_w.is_contiguous() and fc_w.shape[0] <= 64
            and dw_out.is_contiguous()):
        LAUNCHES["head"] += 2
        loss, correct, dfeat, logits = C.head_fwd_bwd(feat, fc_w, fc_b, labels, loss_scale, n_valid, dw_out,
                                                      db_out, accumulate, need_dfeat, ARENA.take(1, 2, feat.device))
        return loss, correct, (dfeat if need_dfeat else None), logits
    _fallback("head_fwd_bwd", str(tuple(feat.shape)))
    return _tb.head_fwd_bwd(feat, fc_w, fc_b, labels, loss_scale, n_valid, dw_out, db_out, accumulate, need_dfeat)


def linear_fwd(x2d, w, b):
    return _tb.linear_fwd(x2d, w, b)


def adam_step(master, grad, m, v, shadow, step_t, lr, b1, b2, eps, grad_scale=1.0, prev=None, zero_grad=False,
              live_blocks=None, diff_out=None, bump=True, max_ctas=0):
    if _dev(master) and master.numel() % 4 == 0 and (shadow is None or shadow.dtype == torch.bfloat16):
        LAUNCHES["adam"] += 2 if bump else 1
        diff = diff_out
        if diff is None and prev is not None:
            diff = torch.empty((), dtype=torch.float32, device=master.device)
        C.adam_step(master, grad, m, v, shadow, step_t, lr, b1, b2, eps, grad_scale, prev, diff, zero_grad, live_blocks,
                    bump, max_ctas)
        return diff
    _fallback("adam_step", "")
    return _tb.adam_step(master, grad, m, v, shadow, step_t, lr, b1, b2, eps, grad_scale, prev, zero_grad, live_blocks,
                         diff_out, bump, max_ctas)


def grad_diff_sq(grad, prev):
    if _dev(grad) and grad.numel() % 4 == 0:
        LAUNCHES["grad_diff"] += 1
        return C.grad_diff_sq(grad, prev)
    return _tb.grad_diff_sq(grad, prev)


def stem_prepare(images, mean, std, dtype):
    if _dev(images) and images.dtype == torch.uint8 and dtype == torch.bfloat16:
        LAUNCHES["u8_normalize"] += 1
        return C.u8_normalize(images, mean, std)
    return _tb.stem_prepare(images, mean, std, dtype)


def stats_update(stats, has_prev, loss, correct, batch, diff_sq):
    if _dev(stats) and loss.dtype == torch.float32 and correct.dtype == torch.float32:
        LAUNCHES["stats"] += 1
        C.stats_update(stats, has_prev, loss.detach(), correct.detach(), float(batch), diff_sq)
        return
    _tb.stats_update(stats, has_prev, loss, correct, batch, diff_sq)
