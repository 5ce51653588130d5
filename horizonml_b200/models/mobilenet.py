"""MobileNetV2 on the framework's op set (reference: ``torchvision.models.mobilenet_v2`` behind the legacy container
entrypoint, train.py:60-68, ``MODEL_TYPE=mobilenet``).

Parameter / buffer names and shapes are torchvision's (``features.N.conv.M…``, ``classifier.1.*``), so state dicts are
interchangeable.  What runs where on a B200 (``ops.set_backend("native")``):

* every **1×1 convolution** (expand, project, the 320→1280 head conv — ~95 % of the model's FLOPs) goes through
  ``ops.conv_bn_act``: the tcgen05 implicit-GEMM kernels (channel counts only need to be multiples of 8: the TMA
  zero-fill rule introduced for tensor-parallel shards) with the BN statistics in the conv epilogue and gradients
  written straight into the flat buckets;
* the **depthwise 3×3 convolutions** go through ``ops.dwconv_bn_act`` → ``csrc/depthwise.cu``: a depthwise conv has
  K = 9 and no cross-channel reuse, so it is a register-resident SIMT stencil (8 channels per thread), not a tensor-core
  GEMM; forward leaves the BN sums in its epilogue, the weight gradient lands in the flat bucket;
* the **3-channel stem** (3×3 / stride 2 → 32) is im2col → the same tcgen05 GEMM (like ResNet's 7×7 stem);
* BatchNorm + **ReLU6** (+ the residual add of the stride-1 blocks) are the ``bn_act`` kernels in their generic
  instantiation (any channel count that is a multiple of 8; activation code 2 clamps at 6 and masks the gradient to
  0 < y < 6);
* the classifier (dropout → avg-pool → FC → CE) is ``ops.head_loss``; dropout on the pooled 1280-vector is a PyTorch op.

On CPU / gloo the same graph runs on the PyTorch-op backend (the oracle the kernels are tested against).  ResNet-18
stays the benchmarked model (every reference trainer uses it); this makes ``MODEL_TYPE=mobilenet`` a real model on the
same engine (flat parameters, fused Adam, fused all-reduce) instead of a torchvision + autocast adapter."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from . import resnet as _resnet
from .resnet import BNP, ConvW, _FUSE_RESADD

# (expand ratio t, output channels c, repeats n, stride s) — torchvision.models.mobilenetv2
_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]


class DWConvW(nn.Module):
    """Depthwise 3×3 weight [C, 1, 3, 3] (torchvision's ``Conv2d(C, C, 3, groups=C)``)."""

    def __init__(self, c: int, stride: int):
        super().__init__()
        w = torch.empty(c, 1, 3, 3)
        nn.init.kaiming_normal_(w, mode="fan_out")
        self.weight = nn.Parameter(w)
        self.stride, self.c = stride, c


def _pw(x, conv: ConvW, bn: BNP, relu6: bool, residual=None, training=True, in_link=None, res_link=None,
        bn_src=None, bn_dst=None):
    """1×1 conv → BN → [+residual] → [ReLU6] on the fused op (native tcgen05 path on GPUs)."""
    return ops.conv_bn_act(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, stride=1, pad=0,
                           relu=2 if relu6 else 0, residual=residual, momentum=bn.momentum, eps=bn.eps,
                           training=training, in_link=in_link, res_link=res_link, bn_src=bn_src, bn_dst=bn_dst)


def _dw(x, conv: nn.Module, bn: BNP, training=True, bn_dst=None, bn_src=None):
    """depthwise 3×3 conv → BN → ReLU6 (csrc/depthwise.cu + the generic bn_act kernels on GPUs)."""
    return ops.dwconv_bn_act(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, stride=conv.stride,
                             act=2, momentum=bn.momentum, eps=bn.eps, training=training, bn_dst=bn_dst, bn_src=bn_src)


def _stem(x, conv: nn.Module, bn: BNP, training=True):
    """dense 3×3 / stride-2 conv over the 3 input channels → BN → ReLU6 (im2col + tcgen05 GEMM on GPUs)."""
    return ops.conv_bn_act(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, stride=conv.stride,
                           pad=1, relu=2, momentum=bn.momentum, eps=bn.eps, training=training)


class _CNA(nn.Sequential):
    """torchvision's Conv2dNormActivation as a parameter container: [0] conv weight holder, [1] BN."""

    def __init__(self, conv: nn.Module, bn: BNP):
        super().__init__(conv, bn)


class InvertedResidual(nn.Module):
    def __init__(self, inp: int, oup: int, stride: int, t: int):
        super().__init__()
        hidden = inp * t
        self.use_res = stride == 1 and inp == oup
        self.expand = t != 1
        layers: List[nn.Module] = []
        if self.expand:
            layers.append(_CNA(ConvW(inp, hidden, 1, 1, 0), BNP(hidden)))
        layers.append(_CNA(DWConvW(hidden, stride), BNP(hidden)))
        layers.append(ConvW(hidden, oup, 1, 1, 0))
        layers.append(BNP(oup))
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        t = self.training
        i = 0
        y = x
        # a residual block's input feeds the expand conv and the skip connection: the skip gradient is added inside the
        # expand conv's dgrad epilogue (ops.GradLink, as in ResNet's BasicBlock) instead of by an accumulation kernel
        link = (ops.GradLink(2) if (_FUSE_RESADD and self.use_res and self.expand and t and torch.is_grad_enabled()
                                    and x.requires_grad) else None)
        # BatchNorm-backward sums handed over by the consuming 1×1 conv's dgrad kernel (ops.BNBackLink, HZ_BN_BWD_IN_DGRAD):
        #   the expand BN <- the depthwise conv's dgrad kernel, the depthwise BN <- the project conv (only consumers);
        #   the previous block's project BN <- this block's expand conv, which sees the complete gradient of x either as
        #   x's only consumer (no skip connection) or by folding the skip share in through `link`
        fuse = _resnet._BN_BWD_IN_DGRAD and t and torch.is_grad_enabled()
        prev = getattr(x, "_hz_bn_back", None) if (fuse and self.expand and (not self.use_res or link is not None)) else None
        exp = ops.BNBackLink() if (fuse and self.expand) else None       # the expand BN <- the depthwise conv's dgrad
        mid = ops.BNBackLink() if fuse else None
        nxt = ops.BNBackLink() if fuse else None
        if self.expand:
            y = _pw(y, self.conv[0][0], self.conv[0][1], True, training=t, in_link=link, bn_src=prev, bn_dst=exp)
            i = 1
        y = _dw(y, self.conv[i][0], self.conv[i][1], t, bn_dst=mid, bn_src=exp)
        out = _pw(y, self.conv[i + 1], self.conv[i + 2], False, residual=x if self.use_res else None, training=t,
                  res_link=link, bn_src=mid, bn_dst=nxt)
        if nxt is not None:
            out._hz_bn_back = nxt
        return out


class _StemConv(nn.Module):
    def __init__(self):
        super().__init__()
        w = torch.empty(32, 3, 3, 3)
        nn.init.kaiming_normal_(w, mode="fan_out")
        self.weight = nn.Parameter(w)
        self.stride = 2


class MobileNetV2(nn.Module):
    def __init__(self, num_classes: int = 10, dropout: float = 0.2):
        super().__init__()
        feats: List[nn.Module] = [_CNA(_StemConv(), BNP(32))]
        inp = 32
        for t, c, n, s in _SETTING:
            for i in range(n):
                feats.append(InvertedResidual(inp, c, s if i == 0 else 1, t))
                inp = c
        feats.append(_CNA(ConvW(inp, 1280, 1, 1, 0), BNP(1280)))
        self.features = nn.Sequential(*feats)
        fc = nn.Linear(1280, num_classes)
        nn.init.normal_(fc.weight, 0, 0.01)
        nn.init.zeros_(fc.bias)
        self.classifier = nn.Sequential(nn.Dropout(dropout), fc)
        self.num_classes, self.dropout = num_classes, dropout
        for m in self.modules():                                  # torchvision init: BN weight 1 / bias 0 (BNP default)
            if isinstance(m, ConvW):
                nn.init.kaiming_normal_(m.weight.data, mode="fan_out")

    def feature_map(self, x):
        t = self.training
        x = x.contiguous(memory_format=torch.channels_last)
        y = _stem(x, self.features[0][0], self.features[0][1], t)
        for blk in list(self.features)[1:-1]:
            y = blk(y)
        prev = getattr(y, "_hz_bn_back", None) if (_resnet._BN_BWD_IN_DGRAD and t and torch.is_grad_enabled()) else None
        return _pw(y, self.features[-1][0], self.features[-1][1], True, training=t, bn_src=prev)

    def _pooled(self, x):
        fm = self.feature_map(x)
        f = fm.to(torch.promote_types(fm.dtype, torch.float32)).mean(dim=(2, 3))
        if self.training and self.dropout > 0:
            f = F.dropout(f, self.dropout, True)
        return f

    def forward(self, x):
        """images → logits [N, num_classes]"""
        f = self._pooled(x)
        fc = self.classifier[1]
        return F.linear(f, fc.weight.to(f.dtype), fc.bias.to(f.dtype))

    def forward_loss(self, x, labels, loss_scale: float = 1.0, stats_out: Optional[dict] = None):
        fc = self.classifier[1]
        if not (self.training and self.dropout > 0):
            # no dropout between pool and FC: the head kernel does avg-pool → FC → CE (+ backward) itself
            return ops.head_loss(self.feature_map(x), fc.weight, fc.bias, labels, loss_scale, n_valid=self.num_classes,
                                 stats_out=stats_out)
        f = self._pooled(x)
        feat = f.to(x.dtype if x.dtype != torch.uint8 else torch.float32).view(f.shape[0], f.shape[1], 1, 1)
        feat = feat.contiguous(memory_format=torch.channels_last)
        return ops.head_loss(feat, fc.weight, fc.bias, labels, loss_scale, n_valid=self.num_classes, stats_out=stats_out)


def mobilenet_v2(num_classes: int = 10, seed: Optional[int] = None, dropout: float = 0.2) -> MobileNetV2:
    if seed is not None:
        st = torch.random.get_rng_state()
        torch.manual_seed(seed)
    m = MobileNetV2(num_classes, dropout)
    if seed is not None:
        torch.random.set_rng_state(st)
    return m
