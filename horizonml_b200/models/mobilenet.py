"""MobileNetV2 — only reachable through the legacy container entrypoint (reference train.py:60-68,
``MODEL_TYPE=mobilenet``).  It is a library model (torchvision graph, cuDNN/ATen kernels under
autocast) adapted to the engine's ``forward_loss`` contract; the hand-written kernel path is
ResNet-18, the model every benchmarked reference script uses."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class LibraryModelAdapter(nn.Module):
    def __init__(self, net: nn.Module, num_classes: int):
        super().__init__()
        self.net, self.num_classes = net, num_classes

    def forward(self, x):
        return self.net(x)

    def forward_loss(self, x, labels, loss_scale: float = 1.0, stats_out: Optional[dict] = None):
        if x.is_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = self.net(x.float())
        else:
            logits = self.net(x.float())
        logits = logits.float()
        loss = F.cross_entropy(logits, labels) * loss_scale
        correct = (logits.argmax(1) == labels).sum().float()
        return loss, correct


def mobilenet_v2(num_classes: int = 10, seed: Optional[int] = None) -> LibraryModelAdapter:
    import torchvision
    if seed is not None:
        st = torch.random.get_rng_state()
        torch.manual_seed(seed)
    net = torchvision.models.mobilenet_v2(weights=None, num_classes=num_classes)
    if seed is not None:
        torch.random.set_rng_state(st)
    return LibraryModelAdapter(net, num_classes)
