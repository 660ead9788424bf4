"""AlexNet (torchvision variant: 11x11/4 stem, 5 convs, 3 FC; 233 MiB of gradients in 16 tensors).

Reachable in the reference through ``getattr(models, args.model)()``
(``pytorch_synthetic_benchmark.py:60``).  Keys match ``torchvision.models.alexnet``.
"""
from __future__ import annotations

import torch.nn as nn

from .. import ops
from .layers import Conv2d, Linear, prepare_input


class _Id(nn.Module):
    def forward(self, x):
        return x


class AlexNet(nn.Module):
    input_size = 224

    def __init__(self, num_classes=1000, dropout=0.5):
        super().__init__()
        self.num_classes, self.p = num_classes, dropout
        # indices follow torchvision's Sequential (ReLU / MaxPool slots kept as identities)
        self.features = nn.Sequential(
            Conv2d(3, 64, 11, 4, 2, bias=True), _Id(), _Id(),
            Conv2d(64, 192, 5, 1, 2, bias=True), _Id(), _Id(),
            Conv2d(192, 384, 3, 1, 1, bias=True), _Id(),
            Conv2d(384, 256, 3, 1, 1, bias=True), _Id(),
            Conv2d(256, 256, 3, 1, 1, bias=True), _Id(), _Id(),
        )
        self.classifier = nn.Sequential(_Id(), Linear(256 * 6 * 6, 4096), _Id(), _Id(), Linear(4096, 4096), _Id(),
                                        Linear(4096, num_classes))

    def forward(self, x):
        x = prepare_input(x)
        f = self.features
        x = ops.max_pool2d(f[0](x, relu=True), 3, 2)
        x = ops.max_pool2d(f[3](x, relu=True), 3, 2)
        x = f[6](x, relu=True)
        x = f[8](x, relu=True)
        x = ops.max_pool2d(f[10](x, relu=True), 3, 2)
        x = x.reshape(x.shape[0], -1)           # logical (c, h, w) order, as torchvision
        c = self.classifier
        x = ops.dropout(x, self.p, self.training)
        x = c[1](x, relu=True)
        x = ops.dropout(x, self.p, self.training)
        x = c[4](x, relu=True)
        return c[6](x)[:, : self.num_classes]


def alexnet(**kw):
    return AlexNet(**kw)
