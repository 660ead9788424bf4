"""Running meters and accuracy helpers.

Parity:
* ``AverageMeter`` — reference ``PyTorch_imagenet/src/imagenet_pytorch_horovod.py:128-146``
  (rank-local running average).
* ``Metric``       — reference ``PyTorch_hvd/src/imagenet_pytorch_horovod.py:239-251``
  (cross-rank averaged; the reference does two blocking CPU allreduces per step, we
  accumulate on device and reduce once per read — SURVEY.md K18/K19).
* ``accuracy``     — reference ``PyTorch_imagenet/...:149-163`` (top-k, percent).
"""
from __future__ import annotations

from typing import Sequence

import torch


class AverageMeter:
    """Stores the last value and the running (count-weighted) average."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._val = 0.0
        self._sum = 0.0
        self._count = 0

    def update(self, val, n: int = 1):
        self._val = float(val)
        self._sum += float(val) * n
        self._count += n

    @property
    def val(self) -> float:
        return self._val

    @property
    def count(self) -> int:
        return self._count

    @property
    def avg(self) -> float:
        return self._sum / max(self._count, 1)


class Metric:
    """Cross-rank averaged metric.

    ``update`` accumulates a (device) scalar without a host sync; ``avg`` performs one
    allreduce(mean) of (sum, n) and one D2H read.
    """

    def __init__(self, name: str, device=None):
        self.name = name
        self._sum = torch.zeros((), dtype=torch.float64, device=device)
        self._n = 0
        self._global = True       # every update so far was already a cross-rank mean (piggy-backed)

    def update(self, val, averaged: bool = False):
        """``averaged=True``: ``val`` is already the cross-rank mean (``FusedSGD.averaged_scalars``)."""
        self._global = self._global and averaged
        if torch.is_tensor(val):
            self._sum += val.detach().to(self._sum.device, torch.float64)
        else:
            self._sum += float(val)
        self._n += 1

    @property
    def avg(self) -> torch.Tensor:
        from ..parallel import dist

        local = (self._sum / max(self._n, 1)).to(torch.float32)
        if self._global and self._n > 0:
            return local.cpu()
        return dist.allreduce(local, average=True, name=self.name).cpu()


def accuracy(output: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1,)):
    """Top-k accuracy in percent for each k (tensors of shape [1])."""
    with torch.no_grad():
        maxk = max(topk)
        bsz = target.size(0)
        _, pred = output.float().topk(maxk, 1, True, True)
        correct = pred.t().eq(target.view(1, -1).expand(maxk, bsz))
        return [correct[:k].reshape(-1).float().sum(0, keepdim=True).mul_(100.0 / bsz) for k in topk]


def top1_accuracy(output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Fraction correct (reference ``PyTorch_hvd/...:222-225``)."""
    pred = output.max(1, keepdim=True)[1]
    return pred.eq(target.view_as(pred)).float().mean()
