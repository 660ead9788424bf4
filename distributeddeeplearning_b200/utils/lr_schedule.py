"""Large-batch learning-rate schedule (Goyal et al., arXiv:1706.02677).

Parity: reference ``PyTorch_imagenet/src/imagenet_pytorch_horovod.py:267-289`` and
``PyTorch_hvd/src/imagenet_pytorch_horovod.py:206-219`` (identical maths):

    e   = epoch + (batch_idx + 1) / len(loader)          (only during warm-up)
    adj = (1/size) * (e * (size-1) / warmup_epochs + 1)   for epoch <  warmup_epochs
        = 1.0 / 0.1 / 0.01 / 0.001                        for epoch < 30 / 60 / 80 / else
    lr  = base_lr * size * adj

i.e. a linear ramp base_lr -> base_lr*size over the warm-up, then step decay.
"""
from __future__ import annotations

import logging


def lr_adjustment(epoch: int, batch_idx: int, batches_per_epoch: int, size: int,
                  warmup_epochs: float) -> float:
    if epoch < warmup_epochs:
        e = epoch + float(batch_idx + 1) / max(batches_per_epoch, 1)
        return 1.0 / size * (e * (size - 1) / warmup_epochs + 1)
    if epoch < 30:
        return 1.0
    if epoch < 60:
        return 1e-1
    if epoch < 80:
        return 1e-2
    return 1e-3


def learning_rate(base_lr: float, epoch: int, batch_idx: int, batches_per_epoch: int, size: int,
                  warmup_epochs: float) -> float:
    return base_lr * size * lr_adjustment(epoch, batch_idx, batches_per_epoch, size, warmup_epochs)


def adjust_learning_rate(optimizer, base_lr: float, warmup_epochs: float, batches_per_epoch: int,
                         epoch: int, batch_idx: int, size: int, log: bool = False) -> float:
    """Set ``lr`` on every param group; returns the lr.  Logs only on change (rank 0 caller)."""
    new_lr = learning_rate(base_lr, epoch, batch_idx, batches_per_epoch, size, warmup_epochs)
    for group in optimizer.param_groups:
        if group["lr"] != new_lr:
            group["lr"] = new_lr
            if log:
                logging.getLogger(__name__).info(f"setting lr to {new_lr}")
    return new_lr
