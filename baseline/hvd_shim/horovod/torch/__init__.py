"""horovod.torch API subset used by the reference scripts, implemented on torch.distributed.

Semantics follow Horovod 0.15: DistributedOptimizer registers a hook per parameter that starts an
asynchronous allreduce as soon as that gradient is accumulated, and step() synchronises all handles,
divides by size (average) and then runs the wrapped optimizer.  Compression.fp16 casts before / after.
Extra (measurement only): every step() records a CUDA event so the caller can device-time the run.
"""
import os

import torch
import torch.distributed as dist

_state = {"init": False, "rank": 0, "local_rank": 0, "size": 1}
STEP_EVENTS = []          # one CUDA event per optimizer.step() (device timing by bench.py)


def init():
    if _state["init"]:
        return
    size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank % torch.cuda.device_count())
    if size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": torch.device("cuda", torch.cuda.current_device())} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=size, **kw)
    _state.update(init=True, rank=rank, local_rank=local_rank, size=size)


def rank():
    return _state["rank"]


def local_rank():
    return _state["local_rank"]


def size():
    return _state["size"]


class _NoneCompressor:
    @staticmethod
    def compress(t):
        return t, None

    @staticmethod
    def decompress(t, ctx):
        return t


class _FP16Compressor:
    @staticmethod
    def compress(t):
        if t.dtype.is_floating_point and t.dtype != torch.float16:
            return t.half(), t.dtype
        return t, None

    @staticmethod
    def decompress(t, ctx):
        return t.to(ctx) if ctx is not None else t


class Compression:
    none = _NoneCompressor
    fp16 = _FP16Compressor


def allreduce(tensor, average=True, name=None):
    if size() == 1:
        return tensor.clone()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = tensor.detach().to(dev, copy=True)
    dist.all_reduce(buf)
    if average:
        buf /= size()
    return buf.to(tensor.device)


def broadcast(tensor, root_rank, name=None):
    if size() == 1:
        return tensor.clone()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = tensor.detach().to(dev, copy=True)
    dist.broadcast(buf, src=root_rank)
    return buf.to(tensor.device)


def broadcast_parameters(params, root_rank):
    if size() == 1:
        return
    items = params.items() if isinstance(params, dict) else params
    for _, p in items:
        if torch.is_tensor(p):
            if p.is_cuda or dist.get_backend() != "nccl":
                dist.broadcast(p.data if hasattr(p, "data") else p, src=root_rank)


def broadcast_optimizer_state(optimizer, root_rank):
    if size() == 1:
        return
    box = [optimizer.state_dict() if rank() == root_rank else None]
    dist.broadcast_object_list(box, src=root_rank)
    if rank() != root_rank and box[0]["state"]:
        optimizer.load_state_dict(box[0])


class _DistributedOptimizer:
    def __init__(self, optimizer, named_parameters=None, compression=Compression.none):
        self._opt = optimizer
        self._compression = compression
        self._handles = {}
        self._hooks = []
        if size() > 1:
            for group in optimizer.param_groups:
                for p in group["params"]:
                    if p.requires_grad:
                        self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook()))

    def _make_hook(self):
        def hook(p):
            wire, ctx = self._compression.compress(p.grad)
            work = dist.all_reduce(wire, async_op=True)
            self._handles[p] = (work, wire, ctx)
        return hook

    @property
    def param_groups(self):
        return self._opt.param_groups

    @property
    def state(self):
        return self._opt.state

    def state_dict(self):
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        return self._opt.load_state_dict(sd)

    def zero_grad(self, *a, **k):
        return self._opt.zero_grad(*a, **k)

    def synchronize(self):
        for p, (work, wire, ctx) in self._handles.items():
            work.wait()
            out = self._compression.decompress(wire, ctx)
            if out is not p.grad:
                p.grad.copy_(out)
            p.grad.div_(size())
        self._handles.clear()

    def step(self, closure=None):
        self.synchronize()
        out = self._opt.step(closure)
        if torch.cuda.is_available():
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            STEP_EVENTS.append(ev)
        return out


def DistributedOptimizer(optimizer, named_parameters=None, compression=Compression.none):
    return _DistributedOptimizer(optimizer, named_parameters, compression)
