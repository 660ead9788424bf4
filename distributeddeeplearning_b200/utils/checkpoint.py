"""Checkpoint / resume.

Parity: ``PyTorch_hvd/src/imagenet_pytorch_horovod.py`` — rank 0 saves ``{'model','optimizer'}``
per epoch to ``checkpoint-{epoch}.pth.tar`` (``:228-235``); at start every rank scans for the
newest file (``:62-67``), rank 0's answer is broadcast (``:71-72``), rank 0 loads (``:135-139``)
and parameters + optimizer state are broadcast (``:143-144``).  ``PyTorch_imagenet`` saves to a
single path every epoch and is broken off rank 0 — fixed here.
Extra state kept for exact resume: epoch, global step, RNG counters.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch

from ..parallel import dist


def find_resume_epoch(checkpoint_format: str, max_epochs: int) -> int:
    """Newest epoch E (1-based count of finished epochs) whose checkpoint file exists, else 0."""
    for try_epoch in range(max_epochs, 0, -1):
        if os.path.exists(checkpoint_format.format(epoch=try_epoch)):
            return try_epoch
    return 0


def agreed_resume_epoch(checkpoint_format: str, max_epochs: int, root_rank: int = 0) -> int:
    local = find_resume_epoch(checkpoint_format, max_epochs)
    t = dist.broadcast(torch.tensor(local), root_rank=root_rank, name="resume_from_epoch")
    return int(t.item())


def save_checkpoint(path: str, model: torch.nn.Module, optimizer, epoch: int = 0, step: int = 0,
                    extra: Optional[Dict[str, Any]] = None, root_rank: int = 0) -> Optional[str]:
    """All ranks must call (sharded optimizer state is gathered collectively); rank 0 writes.  A pending
    communication error (fused-engine barrier timeout) aborts BEFORE anything is written."""
    if optimizer is not None and hasattr(optimizer, "check_errors"):
        optimizer.check_errors()
    opt_state = optimizer.state_dict() if optimizer is not None else None
    if dist.rank() != root_rank:
        return None
    state = {"model": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "optimizer": opt_state,
             "epoch": epoch, "step": step, "extra": extra or {}}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)
    return path


def load_checkpoint(path: str, model: torch.nn.Module, optimizer=None, root_rank: int = 0) -> Dict[str, Any]:
    """Rank 0 reads the file; weights and optimizer state reach the other ranks by broadcast."""
    meta: Dict[str, Any] = {}
    if dist.rank() == root_rank:
        state = torch.load(path, map_location="cpu", weights_only=False)
        model.load_state_dict(state["model"])
        if optimizer is not None and state.get("optimizer") is not None:
            optimizer.load_state_dict(state["optimizer"])
        meta = {"epoch": state.get("epoch", 0), "step": state.get("step", 0), "extra": state.get("extra", {})}
    meta = dist.broadcast_object(meta, root_rank)
    if hasattr(optimizer, "broadcast_parameters"):
        optimizer.broadcast_parameters(root_rank)          # fused engine: broadcast kernel over the arena
        buffers = {k: v for k, v in model.state_dict().items() if not isinstance(v, torch.nn.Parameter)}
        dist.broadcast_parameters({k: v for k, v in model.named_buffers()}, root_rank)
    else:
        dist.broadcast_parameters(model.state_dict(), root_rank)
    if optimizer is not None:
        dist.broadcast_optimizer_state(optimizer, root_rank)
    return meta
