"""Pure-PyTorch implementation of the op set.

Two jobs: (1) the numerical oracle every sm_100a kernel is tested against, and (2) the
CPU/gloo plumbing backend (BASELINE.json config #1).  Signatures are identical to
``native_backend`` so the autograd layer in ``ops.functional`` is backend-agnostic.

All activations are logical NCHW tensors stored channels_last (physically NHWC); weights are
logical [Cout, Cin, R, S] stored channels_last (physically [Cout, R, S, Cin]).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def _cl(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t.contiguous()


def conv_fwd(x, w, stride: int, pad: int, want_stats: bool):
    y = _cl(F.conv2d(x, w.to(x.dtype), None, stride, pad))
    if not want_stats:
        return y, None
    yf = y.float()
    sums = torch.stack([yf.sum(dim=(0, 2, 3)), (yf * yf).sum(dim=(0, 2, 3))])
    return y, sums


def bn_act_fwd(y_raw, sums, gamma, beta, rmean, rvar, momentum: float, eps: float,
               residual, relu: bool, training: bool):
    """BatchNorm (batch statistics) + optional residual add + optional ReLU.

    Returns (out, mean, invstd).  ``sums`` (2×C: Σy, Σy²) may come from the conv epilogue."""
    C = y_raw.shape[1]
    cnt = y_raw.numel() // C
    if training:
        if sums is None:
            yf = y_raw.float()
            sums = torch.stack([yf.sum(dim=(0, 2, 3)), (yf * yf).sum(dim=(0, 2, 3))])
        mean = sums[0] / cnt
        var = (sums[1] / cnt - mean * mean).clamp_min(0.0)
        if rmean is not None:
            with torch.no_grad():
                unbiased = var * (cnt / max(cnt - 1, 1))
                rmean.mul_(1 - momentum).add_(mean, alpha=momentum)
                rvar.mul_(1 - momentum).add_(unbiased, alpha=momentum)
    else:
        mean, var = rmean.float(), rvar.float()
    invstd = torch.rsqrt(var + eps)
    scale = (gamma.float() * invstd).view(1, C, 1, 1)
    shift = (beta.float() - mean * gamma.float() * invstd).view(1, C, 1, 1)
    out = y_raw.float() * scale + shift
    if residual is not None:
        out = out + residual.float()
    if relu:
        out = out.clamp_min(0.0)
        if int(relu) == 2:                 # ReLU6 (MobileNetV2)
            out = out.clamp_max(6.0)
    return _cl(out.to(y_raw.dtype)), mean, invstd


class GradSlot:
    """(fp32 tensor, accumulate?) — where a kernel deposits a parameter gradient in place."""
    __slots__ = ("t", "acc")

    def __init__(self, t, acc):
        self.t, self.acc = t, acc

    def put(self, g):
        g = g.to(self.t.dtype).view_as(self.t)
        self.t.add_(g) if self.acc else self.t.copy_(g)


def bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, relu: bool, has_residual: bool,
               dgamma_slot=None, dbeta_slot=None, sums=None):
    """Backward of bn_act_fwd (training mode). Returns (dy_raw, dgamma, dbeta, dres); when slots are
    given dγ/dβ are also written (or accumulated) into them."""
    C = y_raw.shape[1]
    cnt = y_raw.numel() // C
    g = dout.float()
    if relu:
        mask = (out > 0) if int(relu) != 2 else ((out > 0) & (out < 6))
        g = g * mask.to(g.dtype)
    dres = _cl(g.to(dout.dtype)) if has_residual else None
    xhat = (y_raw.float() - mean.view(1, C, 1, 1)) * invstd.view(1, C, 1, 1)
    if sums is not None:                   # (Σg, Σg·x̂) handed over by conv_dgrad_bnbwd of the consuming layer
        dbeta, dgamma = sums[0].float(), sums[1].float()
    else:
        dbeta = g.sum(dim=(0, 2, 3))
        dgamma = (g * xhat).sum(dim=(0, 2, 3))
    k = (gamma.float() * invstd).view(1, C, 1, 1)
    dy = k * (g - (dbeta / cnt).view(1, C, 1, 1) - xhat * (dgamma / cnt).view(1, C, 1, 1))
    if dgamma_slot is not None:
        dgamma_slot.put(dgamma)
        dbeta_slot.put(dbeta)
    return _cl(dy.to(y_raw.dtype)), dgamma, dbeta, dres


def bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma, relu, dgamma_slot, dbeta_slot, sums, res_yraw, res_mean,
                   res_invstd):
    """Oracle of the native apply-with-residual-BN-sums kernel: bn_act_bwd of a layer with a residual whose producer
    is a BatchNorm without activation, plus that BatchNorm's backward sums [Σ dres, Σ dres·x̂_res] taken from dres
    rounded to its storage dtype.  Returns (dy, dres, res_sums) or None (activation not covered)."""
    if int(relu) not in (0, 1):
        return None
    dy, _, _, dres = bn_act_bwd(dout, out, y_raw, mean, invstd, gamma, relu, True, dgamma_slot, dbeta_slot, sums=sums)
    C = y_raw.shape[1]
    g = dres.float()
    xhat = (res_yraw.float() - res_mean.view(1, C, 1, 1)) * res_invstd.view(1, C, 1, 1)
    return dy, dres, torch.stack([g.sum(dim=(0, 2, 3)), (g * xhat).sum(dim=(0, 2, 3))])


def conv_dgrad(dy, w, x_shape, stride: int, pad: int, addend=None):
    dx, _, _ = torch.ops.aten.convolution_backward(
        dy, dy.new_empty(x_shape), w.to(dy.dtype), None, [stride, stride], [pad, pad], [1, 1], False,
        [0, 0], 1, [True, False, False])
    if addend is not None:
        dx = dx + addend.to(dx.dtype)
    return _cl(dx)


def conv_dgrad_bnbwd(dy, w, x_shape, stride: int, pad: int, addend, bn_out, bn_yraw, bn_mean, bn_invstd, relu):
    """Oracle of the native dgrad-with-BN-backward-sums kernel: (dx, [Σg, Σg·x̂]) with g = dx·[bn_out > 0] and
    x̂ = (bn_yraw − mean)·invstd, the sums taken from dx rounded to its storage dtype (what the separate reduction pass
    would read back)."""
    if int(relu) not in (0, 1, 2):
        return None
    dx = conv_dgrad(dy, w, x_shape, stride, pad, addend)
    C = x_shape[1]
    g = dx.float()
    if int(relu):
        g = g * ((bn_out > 0) if int(relu) == 1 else ((bn_out > 0) & (bn_out < 6))).to(g.dtype)
    xhat = (bn_yraw.float() - bn_mean.view(1, C, 1, 1)) * bn_invstd.view(1, C, 1, 1)
    return dx, torch.stack([g.sum(dim=(0, 2, 3)), (g * xhat).sum(dim=(0, 2, 3))])


def conv_wgrad(dy, x, w_shape, stride: int, pad: int, out_grad: torch.Tensor, accumulate: bool,
               prezeroed: bool = False):
    """dW written (or accumulated) as fp32 into ``out_grad`` — a view of the flat gradient bucket
    with the weight's logical shape / channels_last strides."""
    _, gw, _ = torch.ops.aten.convolution_backward(
        dy, x, dy.new_empty(w_shape), None, [stride, stride], [pad, pad], [1, 1], False,
        [0, 0], 1, [False, True, False])
    gw = gw.float()
    if accumulate:
        out_grad.add_(gw)
    else:
        out_grad.copy_(gw)


# ---- depthwise 3x3 convolution, pad 1 (MobileNetV2): weight [C, 1, 3, 3]
def dwconv_fwd(x, w, stride: int, want_stats: bool):
    y = _cl(F.conv2d(x, w.to(x.dtype), None, stride, 1, 1, x.shape[1]))
    if not want_stats:
        return y, None
    yf = y.float()
    return y, torch.stack([yf.sum(dim=(0, 2, 3)), (yf * yf).sum(dim=(0, 2, 3))])


def dwconv_dgrad(dy, w, x_shape, stride: int):
    dx, _, _ = torch.ops.aten.convolution_backward(
        dy, dy.new_empty(x_shape), w.to(dy.dtype), None, [stride, stride], [1, 1], [1, 1], False,
        [0, 0], x_shape[1], [True, False, False])
    return _cl(dx)


def dwconv_dgrad_bnbwd(dy, w, x_shape, stride: int, bn_out, bn_yraw, bn_mean, bn_invstd, relu):
    """Oracle of the native depthwise-dgrad-with-BN-backward-sums kernel: (dx, [Σg, Σg·x̂])."""
    if int(relu) not in (0, 1, 2):
        return None
    dx = dwconv_dgrad(dy, w, x_shape, stride)
    C = x_shape[1]
    g = dx.float()
    if int(relu):
        g = g * ((bn_out > 0) if int(relu) == 1 else ((bn_out > 0) & (bn_out < 6))).to(g.dtype)
    xhat = (bn_yraw.float() - bn_mean.view(1, C, 1, 1)) * bn_invstd.view(1, C, 1, 1)
    return dx, torch.stack([g.sum(dim=(0, 2, 3)), (g * xhat).sum(dim=(0, 2, 3))])


def dwconv_wgrad(dy, x, stride: int, out_grad: torch.Tensor, accumulate: bool, prezeroed: bool = False):
    c = x.shape[1]
    _, gw, _ = torch.ops.aten.convolution_backward(
        dy, x, dy.new_empty((c, 1, 3, 3)), None, [stride, stride], [1, 1], [1, 1], False,
        [0, 0], c, [False, True, False])
    gw = gw.float().view_as(out_grad)
    if accumulate:
        out_grad.add_(gw)
    else:
        out_grad.copy_(gw)


def maxpool_fwd(x, want_aux: bool = False):
    """3x3 / stride 2 / pad 1.  Returns (y, aux) - aux is whatever backward needs besides dy."""
    if not want_aux:
        return _cl(F.max_pool2d(x, 3, 2, 1)), None
    y, idx = F.max_pool2d(x, 3, 2, 1, return_indices=True)
    return _cl(y), (idx, tuple(x.shape))


def maxpool_bwd(dy, aux):
    idx, x_shape = aux
    dx = torch.ops.aten.max_pool2d_with_indices_backward(
        dy.contiguous(), dy.new_empty(x_shape), [3, 3], [2, 2], [1, 1], [1, 1], False, idx.contiguous())
    return _cl(dx)


def maxpool_bwd_bn(dy, aux, bn_out, bn_yraw, bn_mean, bn_invstd, relu):
    """Oracle of the native max-pool-backward-with-BN-sums kernel: (dx, [Σg, Σg·x̂])."""
    if int(relu) not in (0, 1):
        return None
    dx = maxpool_bwd(dy, aux)
    C = dx.shape[1]
    g = dx.float()
    if int(relu):
        g = g * (bn_out > 0).to(g.dtype)
    xhat = (bn_yraw.float() - bn_mean.view(1, C, 1, 1)) * bn_invstd.view(1, C, 1, 1)
    return dx, torch.stack([g.sum(dim=(0, 2, 3)), (g * xhat).sum(dim=(0, 2, 3))])


def head_fwd_bwd(feat, fc_w, fc_b, labels, loss_scale: float, n_valid: int,
                 dw_out, db_out, accumulate: bool, need_dfeat: bool = True):
    """global-avg-pool → FC → softmax cross-entropy, forward AND backward in one op.

    feat [N,C,H,W]; fc_w [Kpad, C] (rows ≥ n_valid are padding and masked to −inf);
    returns (loss_sum/N * loss_scale as 0-d fp32, correct count 0-d fp32, dfeat, logits[N,Kpad])."""
    N, C, H, W = feat.shape
    pooled = feat.float().mean(dim=(2, 3))
    logits = pooled @ fc_w.float().t()
    if fc_b is not None:
        logits = logits + fc_b.float()
    K = logits.shape[1]
    if n_valid < K:
        mask = torch.arange(K, device=logits.device) >= n_valid
        logits = logits.masked_fill(mask, float("-inf"))
    lse = torch.logsumexp(logits, dim=1)
    picked = logits.gather(1, labels.view(-1, 1)).squeeze(1)
    loss = (lse - picked).mean() * loss_scale
    correct = (logits.argmax(dim=1) == labels).sum().float()
    p = torch.softmax(logits, dim=1)
    p = p.scatter_add(1, labels.view(-1, 1), -torch.ones(N, 1, device=p.device, dtype=p.dtype))
    dlogits = p * (loss_scale / N)
    dw = dlogits.t() @ pooled
    if accumulate:
        dw_out.add_(dw)
        if db_out is not None:
            db_out.add_(dlogits.sum(0))
    else:
        dw_out.copy_(dw)
        if db_out is not None:
            db_out.copy_(dlogits.sum(0))
    dfeat = None
    if need_dfeat:
        dpooled = dlogits @ fc_w.float()
        dfeat = _cl((dpooled / (H * W)).view(N, C, 1, 1).expand(N, C, H, W).to(feat.dtype))
    return loss, correct, dfeat, logits


def linear_fwd(x2d, w, b):
    y = x2d.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    return y


def adam_step(master, grad, m, v, shadow, step_t, lr: float, b1: float, b2: float, eps: float,
              grad_scale: float = 1.0, prev=None, zero_grad: bool = False, live_blocks=None, diff_out=None,
              bump: bool = True, max_ctas: int = 0):
    """Flat fused Adam (torch.optim.Adam semantics, no weight decay / amsgrad).
    ``step_t`` holds the step count (incremented here).  ``prev``: gradient-divergence bookkeeping —
    returns Σ(g−prev)² and sets prev ← g.  ``zero_grad``: g ← 0 afterwards (accumulate-only wgrads)."""
    if bump:                       # first (or only) bucket of this optimizer step
        step_t += 1
        if diff_out is not None:
            diff_out.zero_()
    t = float(step_t.item()) if step_t.device.type == "cpu" else step_t.float()
    diff = diff_out
    if prev is not None:
        d = grad - prev
        if diff_out is not None:
            diff_out += (d * d).sum()
        else:
            diff = (d * d).sum()
        prev.copy_(grad)
    if (master.device.type == "cpu" and grad_scale == 1.0 and master.dtype == torch.float32
            and hasattr(torch, "_fused_adam_")):
        # CPU path (BASELINE config #1, tests): ATen's vectorised multi-threaded fused Adam — one pass over the flat
        # buffers instead of eight elementwise passes (11 M parameters: 2.7 ms vs 42 ms)
        torch._fused_adam_([master], [grad], [m], [v], [], [step_t.view(())], lr=lr, beta1=b1, beta2=b2,
                           weight_decay=0.0, eps=eps, amsgrad=False, maximize=False)
        if shadow is not None:
            shadow.copy_(master)
        if zero_grad:
            grad.zero_()
        return diff
    g = grad if grad_scale == 1.0 else grad * grad_scale
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    denom = (v / bc2).sqrt_().add_(eps)
    master.sub_((m / bc1) / denom * lr)
    if shadow is not None:
        shadow.copy_(master)
    if zero_grad:
        grad.zero_()
    return diff


def grad_diff_sq(grad, prev):
    """Σ (g_t − g_{t−1})², then prev ← g_t  (reference metric, data_parallel_train.py:132-145)."""
    d = (grad - prev)
    out = (d * d).sum()
    prev.copy_(grad)
    return out


def stem_prepare(images, mean: float, std: float, dtype):
    """uint8/fp32 NCHW images → normalised channels_last activations in the compute dtype."""
    x = images.float()
    if images.dtype == torch.uint8:
        x = x / 255.0
    x = (x - mean) / std
    return _cl(x.to(dtype))


def step_begin():
    """Start-of-step hook (the native backend recycles its statistics arena here)."""


def step_end():
    pass


def stats_update(stats, has_prev, loss, correct, batch: float, diff_sq):
    stats[0] += loss.detach().float()
    stats[1] += correct.detach().float()
    stats[2] += batch
    stats[5] += 1
    if diff_sq is not None:
        hp = has_prev.reshape(())
        stats[3] += diff_sq.reshape(()).sqrt() * hp
        stats[4] += hp
        has_prev.fill_(1.0)
        diff_sq.zero_()         # consumed (bucket-wise optimizer passes accumulate into it from zero)
