"""``DistributedOptimizer``: the reference's one parallelism strategy — synchronous data parallelism
with allreduce-averaged gradients — behind Horovod's call shape:

    optimizer = DistributedOptimizer(optim.SGD(model.parameters(), lr=...),
                                     named_parameters=model.named_parameters(),
                                     compression=Compression.fp16)

(reference ``pytorch_synthetic_benchmark.py:66-74``).  Two implementations:

* CUDA + ``torch.optim.SGD``  ->  ``engine.FusedSGD`` (hand-written NVLink kernels; the product)
* anything else (CPU/gloo plumbing mode, other optimizers) -> ``HookedDistributedOptimizer``:
  per-parameter post-accumulate hooks, static buckets, one ``torch.distributed`` allreduce per
  bucket launched as soon as the bucket is complete, ``step()`` waits on the handles — Horovod's
  semantics without its negotiation thread.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as td

from . import dist
from .compression import Compression


class HookedDistributedOptimizer:
    """Wraps any torch optimizer; averages gradients across ranks before ``step()``."""

    def __init__(self, optimizer: torch.optim.Optimizer, named_parameters=None, compression=Compression.none,
                 bucket_mb: float = 25.0, first_bucket_mb: float = 1.0):
        self.optimizer = optimizer
        self.compression = compression
        params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        if named_parameters is not None:
            named = list(named_parameters)
            names = {id(p): n for n, p in named}
            missing = [p for p in params if id(p) not in names]
            if missing and named:
                raise ValueError("named_parameters does not cover every optimized parameter")
        self._params = list(reversed(params))                       # gradient-ready order
        self._buckets: List[List[torch.Tensor]] = []
        cap, cur, used = first_bucket_mb * (1 << 20), [], 0
        for p in self._params:
            nbytes = p.numel() * p.element_size()
            if cur and used + nbytes > cap:
                self._buckets.append(cur)
                cur, used, cap = [], 0, bucket_mb * (1 << 20)
            cur.append(p)
            used += nbytes
        if cur:
            self._buckets.append(cur)
        self._bucket_of = {id(p): b for b, ps in enumerate(self._buckets) for p in ps}
        self._pending = [len(b) for b in self._buckets]
        self._seen = set()
        self._handles: List[Tuple[int, object, torch.Tensor, object]] = []
        self._next = 0
        self._hooks = []
        if dist.is_distributed():
            for p in self._params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(p)))

    # ---- torch.optim surface ----------------------------------------------------------------------
    @property
    def param_groups(self):
        return self.optimizer.param_groups

    @property
    def state(self):
        return self.optimizer.state

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, sd):
        return self.optimizer.load_state_dict(sd)

    def zero_grad(self, set_to_none: bool = True):
        return self.optimizer.zero_grad(set_to_none=set_to_none)

    def add_param_group(self, g):
        raise NotImplementedError("add parameters before wrapping the optimizer")

    # ---- gradient exchange ------------------------------------------------------------------------
    def _make_hook(self, p):
        def hook(_p):
            if id(p) in self._seen:
                return
            self._seen.add(id(p))
            b = self._bucket_of[id(p)]
            self._pending[b] -= 1
            while self._next < len(self._buckets) and self._pending[self._next] == 0:
                self._launch(self._next)
                self._next += 1
        return hook

    def _launch(self, b: int) -> None:
        ps = [p for p in self._buckets[b]]
        grads = [(p.grad if p.grad is not None else torch.zeros_like(p)) for p in ps]
        flat = torch.cat([g.reshape(-1) for g in grads])
        wire, ctx = self.compression.compress(flat)
        if wire.device.type == "cpu" and wire.dtype in (torch.float16, torch.bfloat16):
            wire = wire.float()        # gloo has no 16-bit reduction; precision loss already applied
            wire = wire.to(self.compression.wire_dtype).float() if self.compression.wire_dtype else wire
        work = td.all_reduce(wire, async_op=True)
        self._handles.append((b, work, wire, ctx))

    def synchronize(self) -> None:
        if not dist.is_distributed():
            return
        while self._next < len(self._buckets):       # parameters without gradients this step
            self._launch(self._next)
            self._next += 1
        size = dist.size()
        for b, work, wire, ctx in self._handles:
            work.wait()
            flat = self.compression.decompress(wire, ctx)
            off = 0
            for p in self._buckets[b]:
                n = p.numel()
                g = flat[off:off + n].view_as(p).to(p.dtype) / size
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self._handles.clear()
        self._pending = [len(b) for b in self._buckets]
        self._seen.clear()
        self._next = 0

    def step(self, closure=None):
        self.synchronize()
        return self.optimizer.step(closure)


def DistributedOptimizer(optimizer: torch.optim.Optimizer, named_parameters=None, compression=Compression.none,
                         fused: Optional[bool] = None, **engine_kw):
    """Factory with Horovod's signature.  ``fused=None`` picks the fused NVLink engine whenever the
    wrapped optimizer is plain SGD on CUDA parameters."""
    params = [p for g in optimizer.param_groups for p in g["params"]]
    on_cuda = bool(params) and all(p.is_cuda for p in params)
    is_sgd = type(optimizer) is torch.optim.SGD and len(optimizer.param_groups) == 1
    maximize = bool(optimizer.param_groups[0].get("maximize", False)) if is_sgd else False
    if fused is None:
        fused = on_cuda and is_sgd and not maximize
    if fused:
        if not (on_cuda and is_sgd):
            raise ValueError("the fused engine needs torch.optim.SGD with one param group on CUDA")
        from .engine import FusedSGD

        g = optimizer.param_groups[0]
        comp = compression
        if compression is Compression.fp16:
            comp = Compression.bf16      # NVLS reduces bf16 natively; same 16-bit wire, fp32 range
        return FusedSGD(params, lr=g["lr"], momentum=g["momentum"], dampening=g["dampening"],
                        weight_decay=g["weight_decay"], nesterov=g["nesterov"], compression=comp, **engine_kw)
    return HookedDistributedOptimizer(optimizer, named_parameters, compression)
