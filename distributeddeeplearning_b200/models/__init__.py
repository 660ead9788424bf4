"""Model zoo (same names as ``torchvision.models`` so ``--model <name>`` keeps working).

Parity: the reference resolves ``--model`` with ``getattr(torchvision.models, name)()``
(``pytorch_synthetic_benchmark.py:60``) / ``models.__dict__[model](pretrained=False)``
(``imagenet_pytorch_horovod.py:383``).  North-star zoo: ResNet-50/101/152, VGG-16, Inception-v3,
AlexNet; ResNet-18/34, the other VGG depths, DenseNet and SqueezeNet (the rest of torchvision
0.2.1's families) are here too.
"""
from __future__ import annotations

from .alexnet import AlexNet, alexnet  # noqa: F401
from .densenet import DenseNet, densenet121, densenet161, densenet169, densenet201  # noqa: F401
from .inception import Inception3, inception_v3  # noqa: F401
from .resnet import ResNet, resnet18, resnet34, resnet50, resnet101, resnet152  # noqa: F401
from .squeezenet import SqueezeNet, squeezenet1_0, squeezenet1_1  # noqa: F401
from .vgg import VGG, vgg11, vgg11_bn, vgg13, vgg13_bn, vgg16, vgg16_bn, vgg19, vgg19_bn  # noqa: F401

_REGISTRY = {
    f.__name__: f
    for f in (alexnet, densenet121, densenet161, densenet169, densenet201, squeezenet1_0, squeezenet1_1, inception_v3, resnet18, resnet34, resnet50, resnet101, resnet152, vgg11, vgg11_bn, vgg13,
              vgg13_bn, vgg16, vgg16_bn, vgg19, vgg19_bn)
}


def available_models():
    return sorted(_REGISTRY)


def get_model(name: str, **kw):
    """Build a model by torchvision name (``pretrained`` is accepted and must be False: no network)."""
    if kw.pop("pretrained", False):
        raise ValueError("pretrained weights are not available offline")
    try:
        return _REGISTRY[name](**kw)
    except KeyError:
        raise ValueError(f"unknown model {name!r}; available: {', '.join(available_models())}") from None


def input_size(model_or_name) -> int:
    """Spatial input size (299 for Inception-v3 — the reference hard-codes 224 and breaks)."""
    if isinstance(model_or_name, str):
        return 299 if model_or_name.startswith("inception") else 224
    return int(getattr(model_or_name, "input_size", 224))
