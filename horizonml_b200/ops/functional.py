"""Backend-agnostic autograd layer over the fused op set.

Design points
-------------
* The unit of fusion is **conv → BatchNorm(batch stats) → [+residual] → [ReLU]**
  (``conv_bn_act``): on the native backend the conv kernel's epilogue produces the BN
  partial sums, one elementwise kernel applies BN/residual/ReLU; backward is
  (reduce, apply, dgrad, wgrad) — see SURVEY §2.5 W4-W6.
* Weight gradients are **written in place into ``param.main_grad``** (a view of the flat
  fp32 gradient bucket owned by ``FlatParams``), never materialised as autograd outputs —
  this is what lets wgrad kernels deposit straight into the all-reduce bucket and lets
  the DP reducer launch a bucket's all-reduce the moment its last wgrad is enqueued
  (reference equivalent: DDP reducer hooks fired inside ``loss.backward()``,
  data_parallel_train.py:118).
* Parameters are fp32 masters; when the compute dtype is bf16 the kernels read the
  ``param.shadow`` bf16 copy maintained by the fused Adam kernel.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import torch_backend as _tb

_state = {"backend": "torch", "native": None}
# wgrad kernels depend only on (dy, x): they run on a side stream, off the dgrad critical path
_side = {"enabled": False, "stream": None, "keep": [], "dirty": False}


def enable_side_stream(flag: bool) -> None:
    _side["enabled"] = bool(flag)


def join_side() -> None:
    """Make the current stream wait for everything launched on the side stream (call before the
    optimizer / at the end of backward; also required before a CUDA-graph capture ends)."""
    if _side["dirty"]:
        ev = torch.cuda.Event()
        ev.record(_side["stream"])
        torch.cuda.current_stream().wait_event(ev)
        _side["keep"].clear()
        _side["dirty"] = False


def step_begin(device=None) -> None:
    if _state["backend"] == "native" and (device is None or torch.device(device).type == "cuda"):
        _state["native"].step_begin(device)


def step_end() -> None:
    if _state["backend"] == "native" and _state["native"] is not None:
        _state["native"].step_end()


def set_backend(name: str) -> None:
    if name not in ("torch", "native"):
        raise ValueError(name)
    if name == "native":
        from . import native_backend  # noqa: WPS433 (lazy: needs the CUDA extension)
        _state["native"] = native_backend
    _state["backend"] = name


def get_backend() -> str:
    return _state["backend"]


def _be(t: torch.Tensor):
    if _state["backend"] == "native" and t.is_cuda:
        return _state["native"]
    return _tb


# ----------------------------------------------------------------------------------------------
# parameter helpers
# ----------------------------------------------------------------------------------------------

def compute_weight(p: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """The tensor kernels should read for parameter ``p`` in compute dtype ``dtype``."""
    if dtype == p.dtype:
        return p.detach()
    sh = getattr(p, "shadow", None)
    if sh is not None and sh.dtype == dtype:
        return sh
    return p.detach().to(dtype)


def grad_target(p: torch.Tensor):
    """(fp32 tensor to write dW into, accumulate?) for parameter ``p``."""
    mg = getattr(p, "main_grad", None)
    if mg is not None:
        return mg, bool(getattr(p, "_acc", False))
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.preserve_format)
    return p.grad, True


def grad_written(p: torch.Tensor) -> None:
    if getattr(p, "main_grad", None) is not None:
        p._acc = True
    hook = getattr(p, "_ready_hook", None)
    if hook is not None:
        hook(p)


def _write_vec_grad(p, g):
    tgt, acc = grad_target(p)
    if acc:
        tgt.add_(g.to(tgt.dtype).view_as(tgt))
    else:
        tgt.copy_(g.to(tgt.dtype).view_as(tgt))
    grad_written(p)


# ----------------------------------------------------------------------------------------------
# conv + BN + (residual) + (ReLU)
# ----------------------------------------------------------------------------------------------

class GradLink:
    """Gradient hand-off between the ``n`` autograd nodes that consume ONE tensor (in a BasicBlock: conv1 and the
    identity / downsample branch).  Instead of every node returning its share and autograd launching an add
    kernel, a node that is not the last to run parks its share here and returns None; the next node folds the
    parked tensor into its dgrad epilogue (``addend``); the last one returns the total.  Order independent."""
    __slots__ = ("n", "k", "pending")

    def __init__(self, n: int = 2):
        self.n, self.k, self.pending = n, 0, None

    def take(self):
        a, self.pending = self.pending, None
        return a

    def put(self, g):
        """``g`` already contains whatever ``take()`` returned."""
        self.k += 1
        if self.k >= self.n:
            self.k = 0
            return g
        self.pending = g
        return None


class BNBackLink:
    """Hand-off between a producer ``conv_bn_act`` (passed as ``bn_dst``) and the ONE op that consumes its output
    (passed as ``bn_src``): the consumer's input gradient *is* the producer's BatchNorm upstream gradient, so the
    consumer's dgrad kernel takes the producer's BN-backward sums (Σg, Σg·x̂) in its epilogue and the producer's
    backward skips its reduction pass over dout / out / y_raw (one kernel less per layer pair).  Only used on the
    native backend, for activations none / ReLU, when the consumer's dgrad output is the complete gradient of the
    producer's output: either the consumer is the only one (``single=True``, e.g. conv1 -> conv2 inside a block), or the
    producer's output feeds several ops whose gradients meet in a ``GradLink`` — then the op that runs last (it folds
    the parked shares into its dgrad epilogue) sees the total and takes the sums (``single=False``: a block's bn2 ->
    the next block's conv1 / downsample / skip connection)."""
    __slots__ = ("out", "y_raw", "mean", "invstd", "act", "sums", "single")

    def __init__(self, single: bool = True):
        self.out = self.y_raw = self.mean = self.invstd = self.sums = None
        self.act = 0
        self.single = single

    def clear(self):
        self.out = self.y_raw = self.mean = self.invstd = self.sums = None


class _ConvBNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, rmean, rvar, stride, pad, relu,
                momentum, eps, training, post_conv, post_dgrad, conv_fn, dgrad_fn, in_link=None, res_link=None,
                bn_src=None, bn_dst=None, res_bn_src=None):
        be = _be(x)
        w = compute_weight(weight, x.dtype)
        ctx.post_dgrad = post_dgrad
        ctx.dgrad_fn = dgrad_fn
        ctx.in_link, ctx.res_link = in_link, res_link
        ctx.bn_src, ctx.bn_dst = bn_src, bn_dst
        ctx.res_bn_src = res_bn_src
        ctx.res_ptr = residual.data_ptr() if (residual is not None and res_bn_src is not None) else None
        fused = None
        if conv_fn is None and post_conv is None and training and hasattr(be, "conv_bn_act_fwd"):
            # one kernel: conv + BN statistics + device-wide barrier + normalise / residual / ReLU
            fused = be.conv_bn_act_fwd(x, w, stride, pad, gamma.detach(), beta.detach(), rmean, rvar, momentum, eps,
                                       residual, relu)
        if fused is not None:
            y_raw, out, mean, invstd = fused
        else:
            if conv_fn is not None:        # fused GEMM + collective kernel (tensor parallel): already reduced,
                y_raw, sums = conv_fn(x, w, training)      # BN partial sums of the REDUCED output from its epilogue
            else:
                y_raw, sums = be.conv_fwd(x, w, stride, pad, training and post_conv is None)
                if post_conv is not None:  # row-parallel conv: partial sums → all-reduce before BN
                    y_raw, sums = post_conv(y_raw), None
            out, mean, invstd = be.bn_act_fwd(y_raw, sums, gamma.detach(), beta.detach(), rmean, rvar,
                                              momentum, eps, residual, relu, training)
        ctx.save_for_backward(x, y_raw, out, mean, invstd)
        ctx.params = (weight, gamma, beta)
        ctx.cfg = (stride, pad, relu, residual is not None, training)
        ctx.x_needs_grad = x.requires_grad
        if bn_dst is not None:             # what the consumer's dgrad epilogue needs to take this BN's backward sums
            bn_dst.out, bn_dst.y_raw, bn_dst.mean, bn_dst.invstd, bn_dst.act = out, y_raw, mean, invstd, int(relu)
            bn_dst.sums = None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y_raw, out, mean, invstd = ctx.saved_tensors
        weight, gamma, beta = ctx.params
        stride, pad, relu, has_res, training = ctx.cfg
        if not training:
            raise RuntimeError("conv_bn_act backward requires training=True")
        be = _be(x)
        dout = dout.contiguous(memory_format=torch.channels_last)
        tg, ag = grad_target(gamma)
        tb, ab = grad_target(beta)
        pre_sums = None
        if ctx.bn_dst is not None:
            # the consumer's dgrad kernel produced `dout` AND this BN's backward sums (only if `dout` is its tensor)
            pre_sums = ctx.bn_dst.sums
            ctx.bn_dst.clear()
        rsrc = ctx.res_bn_src
        with_res = None
        if (has_res and rsrc is not None and rsrc.out is not None and rsrc.sums is None and rsrc.single
                and rsrc.act == 0 and ctx.res_link is None and rsrc.out.data_ptr() == ctx.res_ptr
                and hasattr(be, "bn_act_bwd_res")):
            # the residual is the output of a BatchNorm without activation and this op is its only consumer: the apply
            # kernel stores that BatchNorm's upstream gradient (dres) and takes its backward sums on the way
            with_res = be.bn_act_bwd_res(dout, out, y_raw, mean, invstd, gamma.detach(), relu, _tb.GradSlot(tg, ag),
                                         _tb.GradSlot(tb, ab), pre_sums, rsrc.y_raw, rsrc.mean, rsrc.invstd)
        if with_res is not None:
            dy, dres, rsrc.sums = with_res
        elif pre_sums is not None:
            dy, _, _, dres = be.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma.detach(), relu, has_res,
                                           _tb.GradSlot(tg, ag), _tb.GradSlot(tb, ab), sums=pre_sums)
        else:
            dy, _, _, dres = be.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma.detach(), relu, has_res,
                                           _tb.GradSlot(tg, ag), _tb.GradSlot(tb, ab))
        grad_written(gamma)
        grad_written(beta)
        w = compute_weight(weight, x.dtype)
        tgt, acc = grad_target(weight)
        zeroed = bool(getattr(weight, "_zeroed", False))      # buffer known to be all-zero before this step
        if has_res and ctx.res_link is not None:              # identity branch: park dres for conv1's dgrad epilogue
            parked = ctx.res_link.take()
            dres = ctx.res_link.put(dres if parked is None else dres + parked)
        use_side = _side["enabled"] and x.is_cuda and ctx.x_needs_grad
        if use_side:
            if _side["stream"] is None:
                _side["stream"] = torch.cuda.Stream(device=x.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())            # wgrad needs dy only — not the dgrad enqueued next
        # dgrad first: it is the critical path, and every read of this layer's weight is then enqueued before
        # ``grad_written(weight)`` lets a bucket-wise optimizer update that weight
        dx = None
        if ctx.x_needs_grad:
            link = ctx.in_link
            addend = link.take() if link is not None else None
            src = ctx.bn_src
            fused_bn = None
            total_here = (link is None and src is not None and src.single) or (link is not None and link.k == link.n - 1)
            if (src is not None and src.out is not None and src.sums is None and ctx.dgrad_fn is None
                    and ctx.post_dgrad is None and total_here and hasattr(be, "conv_dgrad_bnbwd")
                    and src.out.data_ptr() == x.data_ptr()):
                # x is the producer's BN output and this dgrad's result (with the parked shares of x's other consumers
                # folded into its epilogue) is the complete gradient of x: take the producer's BN-backward sums here
                # (None = shape / activation not covered: plain dgrad below)
                fused_bn = be.conv_dgrad_bnbwd(dy, w, x.shape, stride, pad, addend, src.out, src.y_raw, src.mean,
                                               src.invstd, src.act)
            if fused_bn is not None:
                dx, src.sums = fused_bn
            elif ctx.dgrad_fn is not None:                    # fused dgrad GEMM + all-reduce (+ residual-gradient addend)
                dx = ctx.dgrad_fn(dy, w, addend)
            else:
                if ctx.post_dgrad is None:
                    dx = be.conv_dgrad(dy, w, x.shape, stride, pad, addend)
                else:                                         # column-parallel conv: Σ over shards
                    dx = ctx.post_dgrad(be.conv_dgrad(dy, w, x.shape, stride, pad))
                    if addend is not None:
                        dx = dx + addend
            if link is not None:
                dx = link.put(dx)
        if use_side:
            _side["stream"].wait_event(ev)
            with torch.cuda.stream(_side["stream"]):
                be.conv_wgrad(dy, x, weight.shape, stride, pad, tgt, acc, zeroed)
                grad_written(weight)          # reducer hooks record their events on the side stream
            _side["keep"].append((dy, x))     # keep operands alive until join_side()
            _side["dirty"] = True
        else:
            be.conv_wgrad(dy, x, weight.shape, stride, pad, tgt, acc, zeroed)
            grad_written(weight)
        return (dx, None, None, None, dres) + (None,) * 17


def conv_bn_act(x, weight, gamma, beta, rmean, rvar, stride=1, pad=1, relu=True, residual=None,
                momentum=0.1, eps=1e-5, training=True, post_conv=None, post_dgrad=None,
                conv_fn=None, dgrad_fn=None, in_link=None, res_link=None, bn_src=None, bn_dst=None, res_bn_src=None):
    """``post_conv`` / ``post_dgrad`` are the tensor-parallel reduction points (row-parallel conv
    output, column-parallel conv input-gradient); they take and return a tensor.  ``conv_fn(x, w, want_stats) ->
    (y, sums | None)`` / ``dgrad_fn(dy, w, addend) -> dx`` replace conv + reduction (+ BN statistics pass / residual
    gradient add) by ONE fused GEMM+collective kernel.  ``relu``: False/0 none, True/1 ReLU, 2 ReLU6.
    ``bn_dst`` / ``bn_src``: a ``BNBackLink`` shared by a producer (``bn_dst``) and the single consumer of its output
    (``bn_src``): the consumer's dgrad kernel takes the producer's BatchNorm-backward sums in its epilogue.
    ``res_bn_src``: the ``bn_dst`` link of the op that produced ``residual`` (a BatchNorm without activation whose only
    consumer is this op): this op's BN-backward apply kernel takes that BatchNorm's backward sums."""
    if not training or not torch.is_grad_enabled():
        be = _be(x)
        if conv_fn is not None:
            y_raw, sums = conv_fn(x, compute_weight(weight, x.dtype), training)
        else:
            y_raw, sums = be.conv_fwd(x, compute_weight(weight, x.dtype), stride, pad,
                                      training and post_conv is None)
        if conv_fn is None and post_conv is not None:
            y_raw, sums = post_conv(y_raw), None
        out, _, _ = be.bn_act_fwd(y_raw, sums, gamma.detach(), beta.detach(), rmean, rvar,
                                  momentum, eps, residual, relu, training)
        return out
    return _ConvBNAct.apply(x, weight, gamma, beta, residual, rmean, rvar, stride, pad, relu,
                            momentum, eps, training, post_conv, post_dgrad, conv_fn, dgrad_fn, in_link, res_link,
                            bn_src, bn_dst, res_bn_src)


# ----------------------------------------------------------------------------------------------
# depthwise 3×3 conv (pad 1, stride 1|2) + BN + ReLU6 — MobileNetV2's middle layer (train.py:60-68)
# ----------------------------------------------------------------------------------------------

class _DWConvBNAct(torch.autograd.Function):
    """Same contract as ``_ConvBNAct`` for a depthwise convolution: BN statistics from the conv kernel's epilogue,
    weight gradient written in place into ``weight.main_grad`` (the flat bucket), reducer hooks fired per parameter."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, rmean, rvar, stride, act, momentum, eps, training, bn_dst=None, bn_src=None):
        be = _be(x)
        ctx.bn_src = bn_src
        w = compute_weight(weight, x.dtype)
        y_raw, sums = be.dwconv_fwd(x, w, stride, training)
        out, mean, invstd = be.bn_act_fwd(y_raw, sums, gamma.detach(), beta.detach(), rmean, rvar, momentum, eps, None,
                                          act, training)
        ctx.save_for_backward(x, y_raw, out, mean, invstd)
        ctx.params = (weight, gamma, beta)
        ctx.cfg = (stride, act, training)
        ctx.x_needs_grad = x.requires_grad
        ctx.bn_dst = bn_dst
        if bn_dst is not None:             # the consuming 1×1 conv's dgrad may take this BN's backward sums (BNBackLink)
            bn_dst.out, bn_dst.y_raw, bn_dst.mean, bn_dst.invstd, bn_dst.act = out, y_raw, mean, invstd, int(act)
            bn_dst.sums = None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y_raw, out, mean, invstd = ctx.saved_tensors
        weight, gamma, beta = ctx.params
        stride, act, training = ctx.cfg
        if not training:
            raise RuntimeError("dwconv_bn_act backward requires training=True")
        be = _be(x)
        dout = dout.contiguous(memory_format=torch.channels_last)
        tg, ag = grad_target(gamma)
        tb, ab = grad_target(beta)
        pre_sums = None
        if ctx.bn_dst is not None:
            pre_sums = ctx.bn_dst.sums
            ctx.bn_dst.clear()
        if pre_sums is not None:
            dy, _, _, _ = be.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma.detach(), act, False,
                                        _tb.GradSlot(tg, ag), _tb.GradSlot(tb, ab), sums=pre_sums)
        else:
            dy, _, _, _ = be.bn_act_bwd(dout, out, y_raw, mean, invstd, gamma.detach(), act, False,
                                        _tb.GradSlot(tg, ag), _tb.GradSlot(tb, ab))
        grad_written(gamma)
        grad_written(beta)
        dx = None
        if ctx.x_needs_grad:
            src = ctx.bn_src
            fused = None
            if (src is not None and src.out is not None and src.sums is None and src.single
                    and hasattr(be, "dwconv_dgrad_bnbwd") and src.out.data_ptr() == x.data_ptr()):
                # x is the producing layer's BN output and this depthwise conv its only consumer (BNBackLink)
                fused = be.dwconv_dgrad_bnbwd(dy, compute_weight(weight, x.dtype), x.shape, stride, src.out, src.y_raw,
                                              src.mean, src.invstd, src.act)
            if fused is not None:
                dx, src.sums = fused
            else:
                dx = be.dwconv_dgrad(dy, compute_weight(weight, x.dtype), x.shape, stride)
        tgt, acc = grad_target(weight)
        be.dwconv_wgrad(dy, x, stride, tgt, acc, bool(getattr(weight, "_zeroed", False)))
        grad_written(weight)
        return (dx,) + (None,) * 12


def dwconv_bn_act(x, weight, gamma, beta, rmean, rvar, stride=1, act=2, momentum=0.1, eps=1e-5, training=True,
                  bn_dst=None, bn_src=None):
    """Depthwise 3×3 conv (``weight`` [C,1,3,3], pad 1) → BatchNorm → activation (0 none | 1 ReLU | 2 ReLU6)."""
    if not training or not torch.is_grad_enabled():
        be = _be(x)
        y_raw, sums = be.dwconv_fwd(x, compute_weight(weight, x.dtype), stride, training)
        out, _, _ = be.bn_act_fwd(y_raw, sums, gamma.detach(), beta.detach(), rmean, rvar, momentum, eps, None, act,
                                  training)
        return out
    return _DWConvBNAct.apply(x, weight, gamma, beta, rmean, rvar, int(stride), int(act), momentum, eps, training, bn_dst,
                              bn_src)


# ----------------------------------------------------------------------------------------------
# max-pool 3×3 / stride 2 / pad 1 (the stem pool; index-free recompute backward)
# ----------------------------------------------------------------------------------------------

class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bn_src=None):
        ctx.be = _be(x)
        ctx.bn_src = bn_src
        ctx.x_ptr = x.data_ptr()
        y, ctx.aux = ctx.be.maxpool_fwd(x, True)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous(memory_format=torch.channels_last)
        src = ctx.bn_src
        if (src is not None and src.out is not None and src.sums is None and src.single and hasattr(ctx.be, "maxpool_bwd_bn")
                and src.out.data_ptr() == ctx.x_ptr):
            # the pool is the only consumer of the producing layer's BN output: take that BatchNorm's backward sums here
            fused = ctx.be.maxpool_bwd_bn(dy, ctx.aux, src.out, src.y_raw, src.mean, src.invstd, src.act)
            if fused is not None:
                dx, src.sums = fused
                return dx, None
        return ctx.be.maxpool_bwd(dy, ctx.aux), None


def maxpool3x3s2(x, bn_src=None):
    """``bn_src``: the ``BNBackLink`` of the conv_bn_act that produced ``x`` when this pool is its only consumer."""
    if not x.requires_grad or not torch.is_grad_enabled():
        return _be(x).maxpool_fwd(x, False)[0]
    return _MaxPool.apply(x, bn_src)


# ----------------------------------------------------------------------------------------------
# classifier head: avg-pool → FC → softmax-CE (+ accuracy), forward and backward in one kernel
# ----------------------------------------------------------------------------------------------

class _HeadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, fc_w, fc_b, labels, loss_scale, n_valid, stats_out):
        be = _be(feat)
        tw, accw = grad_target(fc_w)
        tb, accb = grad_target(fc_b) if fc_b is not None else (None, False)
        w = fc_w.detach()
        loss, correct, dfeat, logits = be.head_fwd_bwd(
            feat, w, fc_b.detach() if fc_b is not None else None, labels, loss_scale, n_valid,
            tw, tb, accw, feat.requires_grad)
        ctx.params = (fc_w, fc_b)
        ctx.save_for_backward(dfeat if dfeat is not None else torch.empty(0))
        ctx.has_dfeat = dfeat is not None
        if stats_out is not None:
            stats_out["correct"] = correct
            stats_out["logits"] = logits
        ctx.mark_non_differentiable(correct)
        ctx.set_materialize_grads(False)          # no zero-fill kernel for the (unused) gradient of ``correct``
        return loss, correct

    @staticmethod
    def backward(ctx, dloss, _dcorrect):
        # gradients were produced in forward with d(loss)=1 (loss_scale already folded in)
        fc_w, fc_b = ctx.params
        grad_written(fc_w)
        if fc_b is not None:
            grad_written(fc_b)
        (dfeat,) = ctx.saved_tensors
        return (dfeat if ctx.has_dfeat else None), None, None, None, None, None, None


def head_loss(feat, fc_w, fc_b, labels, loss_scale: float = 1.0, n_valid: Optional[int] = None,
              stats_out: Optional[dict] = None):
    """Returns (loss, correct_count).  ``loss.backward()`` must be seeded with 1 (the default)."""
    n_valid = fc_w.shape[0] if n_valid is None else n_valid
    return _HeadLoss.apply(feat, fc_w, fc_b, labels, float(loss_scale), int(n_valid), stats_out)


_ROOT_GRAD = {}


def backward(loss: torch.Tensor) -> None:
    """``loss.backward()`` seeded with a cached constant 1 — autograd's default root gradient is a fresh
    ``ones_like`` (a fill kernel on the critical path of every step, and not a PDL-aware one)."""
    key = (loss.device, loss.dtype, tuple(loss.shape))
    one = _ROOT_GRAD.get(key)
    if one is None:
        if loss.is_cuda and torch.cuda.is_current_stream_capturing():
            loss.backward()               # never allocate the cache inside a capture (private pool memory)
            return
        one = torch.ones(loss.shape, dtype=loss.dtype, device=loss.device)
        _ROOT_GRAD[key] = one
    torch.autograd.backward(loss, grad_tensors=one)


def head_logits(feat, fc_w, fc_b):
    """Inference-only logits (avg-pool + FC)."""
    pooled = feat.float().mean(dim=(2, 3))
    return _be(feat).linear_fwd(pooled, fc_w.detach(), fc_b.detach() if fc_b is not None else None)


# ----------------------------------------------------------------------------------------------
# small utilities shared by trainers
# ----------------------------------------------------------------------------------------------

def adam_step(master, grad, m, v, shadow, step_t, lr, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0,
              prev=None, zero_grad=False, live_blocks=None, diff_out=None, bump=True, max_ctas=0):
    """Returns Σ(g−prev)² (0-d tensor) when ``prev`` is given, else None.  ``live_blocks``: visit only these
    64-element blocks (all other parameters provably never receive a gradient).  Bucket-wise use: pass slices,
    a shared ``diff_out`` accumulator, and ``bump=True`` only for the first bucket of the step (it advances the
    step counter and clears the accumulator)."""
    return _be(master).adam_step(master, grad, m, v, shadow, step_t, lr, b1, b2, eps, grad_scale, prev, zero_grad,
                                 live_blocks, diff_out, bump, max_ctas)


def grad_diff_sq(grad, prev):
    return _be(grad).grad_diff_sq(grad, prev)


def stats_update(stats, has_prev, loss, correct, batch, diff_sq=None):
    _be(stats).stats_update(stats, has_prev, loss, correct, batch, diff_sq)


def stem_prepare(images, mean=0.5, std=0.5, dtype=torch.float32):
    return _be(images).stem_prepare(images, mean, std, dtype)
