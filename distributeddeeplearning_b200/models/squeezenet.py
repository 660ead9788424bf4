"""SqueezeNet 1.0 / 1.1 (keys match ``torchvision.models.squeezenet1_*``).

Part of the torchvision 0.2.1 zoo the reference's ``--model`` flag reaches.  Every layer is
conv + bias + ReLU (fused in the conv epilogue); Fire modules concatenate their two expand branches with the native
channel-concat kernel; the classifier is a 1x1 conv to ``num_classes`` followed by global average pooling.
"""
from __future__ import annotations

import torch.nn as nn

from .. import ops
from .layers import Conv2d, prepare_input


class Fire(nn.Module):
    def __init__(self, cin, squeeze, e1, e3):
        super().__init__()
        self.squeeze = Conv2d(cin, squeeze, 1, bias=True)
        self.expand1x1 = Conv2d(squeeze, e1, 1, bias=True)
        self.expand3x3 = Conv2d(squeeze, e3, 3, padding=1, bias=True)

    def forward(self, x):
        s = self.squeeze(x, relu=True)
        return ops.concat_channels([self.expand1x1(s, relu=True), self.expand3x3(s, relu=True)])


class _Pool(nn.Module):
    def forward(self, x):
        return ops.max_pool2d(x, 3, 2, 0, ceil_mode=True)


class _Relu(nn.Module):
    """Placeholder keeping torchvision's Sequential indices; the ReLU itself is fused into the preceding conv."""

    def forward(self, x):
        return x


class _Drop(nn.Module):
    def __init__(self, p):
        super().__init__()
        self.p = p

    def forward(self, x):
        return ops.dropout(x, self.p, self.training)


class SqueezeNet(nn.Module):
    input_size = 224

    def __init__(self, version="1_0", num_classes=1000, dropout=0.5):
        super().__init__()
        self.num_classes = num_classes
        if version == "1_0":
            layers = [Conv2d(3, 96, 7, stride=2, bias=True), _Relu(), _Pool(), Fire(96, 16, 64, 64), Fire(128, 16, 64, 64),
                      Fire(128, 32, 128, 128), _Pool(), Fire(256, 32, 128, 128), Fire(256, 48, 192, 192),
                      Fire(384, 48, 192, 192), Fire(384, 64, 256, 256), _Pool(), Fire(512, 64, 256, 256)]
        elif version == "1_1":
            layers = [Conv2d(3, 64, 3, stride=2, bias=True), _Relu(), _Pool(), Fire(64, 16, 64, 64), Fire(128, 16, 64, 64),
                      _Pool(), Fire(128, 32, 128, 128), Fire(256, 32, 128, 128), _Pool(), Fire(256, 48, 192, 192),
                      Fire(384, 48, 192, 192), Fire(384, 64, 256, 256), Fire(512, 64, 256, 256)]
        else:
            raise ValueError(f"unsupported SqueezeNet version {version!r}")
        self.features = nn.Sequential(*layers)
        final = Conv2d(512, num_classes, 1, bias=True)
        self.classifier = nn.Sequential(_Drop(dropout), final, _Relu())
        for m in self.modules():
            if isinstance(m, Conv2d):
                if m is final:
                    nn.init.normal_(m.weight, mean=0.0, std=0.01)
                else:
                    nn.init.kaiming_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = prepare_input(x)
        for m in self.features:
            x = m(x, relu=True) if isinstance(m, Conv2d) else m(x)
        x = self.classifier[0](x)
        x = self.classifier[1](x, relu=True)
        return ops.global_avg_pool(x)


def squeezenet1_0(**kw):
    return SqueezeNet("1_0", **kw)


def squeezenet1_1(**kw):
    return SqueezeNet("1_1", **kw)
