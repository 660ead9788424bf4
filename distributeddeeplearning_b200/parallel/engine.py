"""Fused data-parallel SGD engine: static buckets in a symmetric arena + ONE kernel per bucket that
does allreduce, averaging, (bf16 wire cast,) SGD-momentum update, bf16 weight refresh and gradient
clear over NVLink peer / NVLS multicast memory.

This is the product path for the reference's
``hvd.DistributedOptimizer(optim.SGD(...), named_parameters, compression)`` + ``optimizer.step()``
(``pytorch_synthetic_benchmark.py:66-74,92``; ``imagenet_pytorch_horovod.py:395-405,186``;
``PyTorch_hvd/...:122-131,164``).  Differences by design:

* no runtime negotiation: the bucket plan is a pure function of the parameter list
  (``_C.plan_buckets``), identical on every rank, cross-checked once by hash;
* gradients are produced IN the arena (wgrad kernels ``red.add`` into their bucket slot; generic
  autograd gradients accumulate in place because ``param.grad`` aliases the slot);
* a bucket's kernel is launched on a side stream the moment its last gradient is ready
  (CUDA event), overlapping the rest of backward;
* ``step()`` only joins the side stream; ``zero_grad()`` is a no-op (the kernel clears);
* momentum is sharded by construction (rank r updates slice r of every bucket).

With world_size == 1 the same class runs the local fused SGD kernel per bucket.
"""
from __future__ import annotations

import os
import weakref
import secrets
from typing import Dict, List, Optional, Tuple

import torch

from .. import _ext
from . import dist
from .compression import Compression, Compressor

CL = torch.channels_last
_ALIGN = 4096


class _DevMem:
    """Expose a raw device range to torch through ``__cuda_array_interface__`` (zero copy)."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._owner = owner


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class SymmetricArena:
    """One symmetric allocation per rank carved into named regions (byte offsets identical everywhere)."""

    def __init__(self, regions: List[Tuple[str, int]], device: torch.device, want_multicast: Optional[bool] = None,
                 session: Optional[str] = None, local: bool = False):
        self.C = _ext.load()
        self.rank, self.world = (0, 1) if local else (dist.rank(), dist.size())
        self.device = device
        self.offsets: Dict[str, int] = {}
        cur = 0
        for name, nbytes in regions:
            self.offsets[name] = cur
            cur += _round_up(max(nbytes, 16), _ALIGN)
        self.nbytes = cur
        self.native = None
        self.mc_ptr = 0
        self._regions: Dict[str, torch.Tensor] = {}
        if self.world == 1:
            # no peers: every region is its own torch allocation (own autograd version counter)
            self._buf = None
            for name, nbytes in regions:
                self._regions[name] = torch.zeros(_round_up(max(nbytes, 16), _ALIGN), dtype=torch.uint8, device=device)
            self.peer_ptrs = [0]
        else:
            torch.cuda.synchronize(device)
            dev_index = device.index if device.index is not None else torch.cuda.current_device()
            self.native = self.C.SymmArena(self.rank, self.world, dev_index, self.nbytes)
            self.native.alloc()
            tag = session or dist.broadcast_object(f"ddl{os.getpid()}-{secrets.token_hex(12)}", 0)   # unguessable socket name
            SymmetricArena._seq = getattr(SymmetricArena, "_seq", 0) + 1
            tag = f"{tag}-{SymmetricArena._seq}"
            self.native.exchange(tag + "-x", 60000)
            self.peer_ptrs = list(self.native.peer_ptrs)
            if want_multicast is None:
                want_multicast = os.environ.get("DDL_DISABLE_MULTICAST", "0") != "1"
            ok = 1.0 if (want_multicast and self.native.multicast_supported()) else 0.0
            ok = dist.allreduce_scalar(ok, op="min")
            if ok > 0:
                created = 1.0 if self.native.mc_create(tag + "-m", 60000) else 0.0
                created = dist.allreduce_scalar(created, op="min")   # also the barrier before bind
                if created > 0:
                    bound = 1.0 if self.native.mc_bind() else 0.0
                    bound = dist.allreduce_scalar(bound, op="min")
                    if bound > 0:
                        self.mc_ptr = int(self.native.mc_ptr)
            dist.barrier()
            # one torch base tensor PER REGION (separate autograd version counters: in-place gradient
            # accumulation into the grad region must not invalidate saved views of the weight regions)
            base = self.peer_ptrs[self.rank]
            ends = sorted(self.offsets.values()) + [self.nbytes]
            for name, off in self.offsets.items():
                nb = ends[ends.index(off) + 1] - off
                self._regions[name] = torch.as_tensor(_DevMem(base + off, nb, self.native), device=device)

    @property
    def has_multicast(self) -> bool:
        return self.mc_ptr != 0

    def region(self, name: str, dtype: torch.dtype, numel: int) -> torch.Tensor:
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self._regions[name][:nbytes].view(dtype)


def _param_view(flat: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """View of ``flat`` with p's logical shape and p's memory layout (KRSC for channels_last 4-D)."""
    if p.dim() == 4 and p.is_contiguous(memory_format=CL) and not p.is_contiguous():
        co, ci, r, s = p.shape
        return flat.view(co, r, s, ci).permute(0, 3, 1, 2)
    return flat.view(p.shape)


def _bf16_operand_view(flat: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    if p.dim() == 4:
        return flat.view(p.shape[0], -1)
    return flat.view(p.shape)


_NATIVE_HOOKS_DEFAULT = "1"


class FusedSGD(torch.optim.Optimizer):
    """SGD(+momentum, +weight decay, +nesterov) fused with the data-parallel gradient allreduce."""

    def __init__(self, params, lr: float = 0.01, momentum: float = 0.0, dampening: float = 0.0,
                 weight_decay: float = 0.0, nesterov: bool = False, compression: type = Compression.none,
                 first_bucket_mb: float = 1.0, bucket_mb: float = 16.0, overlap: bool = True,
                 comm_blocks: int = 32, use_multicast: Optional[bool] = None, timeout_s: float = 30.0,
                 broadcast_root: Optional[int] = 0, debug: Optional[bool] = None, oneshot_kb: float = 0.0,
                 local: bool = False):
        """``local=True`` builds a single-rank engine inside a multi-rank job (no peers, no averaging): the oracle of
        ``selfcheck.step_equivalence``."""
        named = list(params)
        if named and isinstance(named[0], tuple):
            plist = [p for _, p in named]
        else:
            plist = named
        plist = [p for p in plist if p.requires_grad]
        if not plist:
            raise ValueError("FusedSGD: no trainable parameters")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(plist, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedSGD supports a single parameter group")
        dev = plist[0].device
        if dev.type != "cuda":
            raise ValueError("FusedSGD needs CUDA parameters (use DistributedOptimizer on CPU)")
        if any(p.dtype != torch.float32 for p in plist):
            raise ValueError("FusedSGD keeps fp32 master weights; parameters must be fp32")
        self.C = _ext.load()
        self.device = dev
        self.world, self.rank = (1, 0) if local else (dist.size(), dist.rank())
        self.wire_bf16 = compression is not Compression.none and compression is not Compressor
        self.compression = compression
        self.overlap = overlap
        self.comm_blocks = int(os.environ.get("DDL_COMM_BLOCKS", comm_blocks))      # env: tuning hook (A/B runs)
        # one-shot buckets (every rank reduces everything, no broadcast phase) always need their own closing barrier;
        # since the two-shot kernels no longer end with one, two-shot wins at every size and one-shot is opt-in
        self.oneshot_bytes = int(float(os.environ.get("DDL_ONESHOT_KB", oneshot_kb)) * 1024)
        # debug mode: after every step verify the protocol's invariants and poison the wire staging
        self.debug = bool(int(os.environ.get("DDL_COMM_DEBUG", "0"))) if debug is None else bool(debug)
        self._sms = torch.cuda.get_device_properties(dev).multi_processor_count

        # ---- static plan: parameters in gradient-ready (reverse registration) order -----------
        self.params: List[torch.Tensor] = list(reversed(plist))
        numels = [p.numel() for p in self.params]
        bucket_mb = float(os.environ.get("DDL_BUCKET_MB", bucket_mb))                 # env: tuning hook (A/B runs)
        plan = self.C.plan_buckets(numels, max(2048, int(first_bucket_mb * (1 << 20) / 4)),
                                   max(2048, int(bucket_mb * (1 << 20) / 4)), 64, 2048)
        self.plan = plan
        h = float(plan["hash"] % (1 << 52))
        if dist.is_distributed() and not local and (dist.allreduce_scalar(h, op="max") != h or dist.allreduce_scalar(h, op="min") != h):
            raise RuntimeError("bucket plan differs across ranks (models are not identical)")
        T = int(plan["total_elems"])
        self.total_elems = T
        self.num_buckets = len(plan["bucket_start"])

        # ---- arena -----------------------------------------------------------------------------
        regions = [("flags", int(self.C.SIGNAL_PAD_BYTES)), ("grad", T * 4), ("weight", T * 4), ("wbf16", T * 2)]
        if self.wire_bf16:
            regions.append(("stage", T * 2))
        regions.append(("scalars", int(self.C.SCALAR_SLOTS) * 4))
        self.arena = SymmetricArena(regions, dev, use_multicast, local=local)
        self.use_mc = self.arena.has_multicast
        self.G = self.arena.region("grad", torch.float32, T)
        self.W = self.arena.region("weight", torch.float32, T)
        self.Wb = self.arena.region("wbf16", torch.bfloat16, T)
        self.M = torch.zeros(T, dtype=torch.float32, device=dev)
        self._scalars_in = self.arena.region("scalars", torch.float32, int(self.C.SCALAR_SLOTS))
        self._scalars_out = torch.zeros(int(self.C.SCALAR_SLOTS), dtype=torch.float32, device=dev)
        self._scalars_pending = 0
        self._scalars_last = 0
        self._epochs = torch.zeros(int(self.C.COMM_CHANNELS) * int(self.C.MAX_COMM_BLOCKS), dtype=torch.int32, device=dev)
        self._error = torch.zeros(1, dtype=torch.int32, device=dev)
        self._hyper_dev = torch.zeros(int(self.C.SGD_HYPER_BYTES), dtype=torch.uint8, device=dev)
        self._hyper_host = torch.zeros(int(self.C.SGD_HYPER_BYTES), dtype=torch.uint8).pin_memory()
        self._hyper_blob = None
        self._hyper_event = None
        off = self.arena.offsets
        self.ctx = self.C.CommCtx(self.arena.peer_ptrs, self.arena.mc_ptr, self.rank, self.world, off["flags"],
                                  off["grad"], off["weight"], off["wbf16"], off.get("stage", 0),
                                  self._epochs.data_ptr(), self._error.data_ptr(), int(timeout_s * 1e9),
                                  int(os.environ.get("DDL_COMM_SKEW_NS", "0")))

        # ---- move parameters into the arena ----------------------------------------------------
        self._index = {}
        with torch.no_grad():
            for i, p in enumerate(self.params):
                o, n = int(plan["param_offset"][i]), p.numel()
                wv = _param_view(self.W[o:o + n], p)
                wv.copy_(p.data)
                p.data = wv
                p.grad = _param_view(self.G[o:o + n], p)
                p._ddl_bf16 = _bf16_operand_view(self.Wb[o:o + n], p)
                self._index[id(p)] = i
        self._pending = list(plan["bucket_param_count"])
        self._ready_seen = [False] * len(self.params)
        self._next_bucket = 0
        self._hyper_uploaded = False
        self._first_step = True
        self._comm_stream = torch.cuda.Stream(device=dev, priority=-1) if overlap else None
        self._steps = 0
        self._hold = False
        # ---- gradient-ready hooks ---------------------------------------------------------------------
        # native (default): the per-parameter hook is a C++ callable of the StepLauncher (csrc/runtime/step_launcher.h):
        # bucket accounting, stream joins and the bucket kernel launch happen without a Python frame; Python is called
        # back once per step (current stream, hyper-parameter upload).  DDL_NATIVE_HOOKS=0: the same logic in Python.
        self._launcher = None
        self._launches_seen = 0
        if os.environ.get("DDL_NATIVE_HOOKS", _NATIVE_HOOKS_DEFAULT) != "0":
            from ..ops import functional as _F

            self._launcher = self.C.StepLauncher(
                [int(v) for v in plan["param_bucket"]], [int(v) for v in plan["bucket_param_count"]],
                [int(v) for v in plan["bucket_start"]], [int(v) for v in plan["bucket_numel"]], self.ctx, self.world,
                self.W.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.Wb.data_ptr(), self._hyper_dev.data_ptr(),
                self.comm_blocks, self._sms, bool(self.use_mc), bool(self.wire_bf16), self.oneshot_bytes,
                self.arena.offsets["scalars"], self._scalars_out.data_ptr(),
                self._comm_stream.cuda_stream if self._comm_stream is not None else 0,
                lambda d=dev: torch.cuda.current_stream(d).cuda_stream,
                lambda handle, r=weakref.ref(self): r()._upload_hyper_handle(handle))      # no C++-held cycle
            _F.use_native_wgrad_join(True)
        self._hook_handles = []
        self._register_hooks()
        if broadcast_root is not None and self.world > 1:
            self.broadcast_parameters(broadcast_root)
        self.refresh_bf16()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------------------------------------
    def _lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    def refresh_hyper_host(self) -> None:
        """Pack the current hyper-parameters (lr schedule!) into the pinned staging buffer.  The H2D copy that follows
        in every step is a memcpy node when the step is replayed from a CUDA graph, so calling this before
        ``graph.replay()`` is all a captured step needs to follow an LR schedule.

        The blob is rewritten only when its contents change, and only after the device has consumed the previous
        contents (``mark_hyper_consumed`` records an event behind the last enqueued copy): the host may run several
        steps ahead of the GPU, and a queued step must not pick up a later step's lr / first_step."""
        g = self.param_groups[0]
        blob = self.C.pack_sgd_hyper(float(g["lr"]), float(g["momentum"]), float(g["dampening"]),
                                     float(g["weight_decay"]), 1.0 / self.world, bool(g["nesterov"]),
                                     bool(self._first_step))
        if blob == self._hyper_blob:
            return
        if self._hyper_event is not None:
            self._hyper_event.synchronize()
        self._hyper_host.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        self._hyper_blob = blob

    def mark_hyper_consumed(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Record that every copy of the pinned hyper blob enqueued so far precedes this point of ``stream``."""
        if self._hyper_event is None:
            self._hyper_event = torch.cuda.Event()
        self._hyper_event.record(stream if stream is not None else torch.cuda.current_stream(self.device))

    def _upload_hyper(self, stream: torch.cuda.Stream) -> None:
        self.refresh_hyper_host()
        with torch.cuda.stream(stream):
            self._hyper_dev.copy_(self._hyper_host, non_blocking=True)
        if not torch.cuda.is_current_stream_capturing():
            self.mark_hyper_consumed(stream)

    def _register_hooks(self) -> None:
        for h in self._hook_handles:
            h.remove()
        self._hook_handles = []
        for i, p in enumerate(self.params):
            if self._launcher is not None:
                hook = self._launcher.hook(i)
                p._ddl_ready = hook
                self._hook_handles.append(p.register_post_accumulate_grad_hook(hook))
            else:
                p._ddl_ready = (lambda idx=i: self._on_ready(idx))
                self._hook_handles.append(p.register_post_accumulate_grad_hook(lambda _p, idx=i: self._on_ready(idx)))

    def use_python_hooks(self) -> None:
        """Switch this optimizer to the Python hook path (between steps): the timing tools that wrap ``_launch_bucket``
        (``workloads.benchmark.phase_times``, ``tools/comm_timeline.py``) need a Python frame per bucket launch."""
        if self._launcher is None:
            return
        from ..ops import functional as _F

        self._launcher = None
        _F.use_native_wgrad_join(False)
        self._register_hooks()

    @property
    def _hold_buckets(self) -> bool:
        return self._hold

    @_hold_buckets.setter
    def _hold_buckets(self, on: bool) -> None:          # selfcheck: keep the gradients in the arena until step()
        self._hold = bool(on)
        if self._launcher is not None:
            self._launcher.set_hold(bool(on))

    def _upload_hyper_handle(self, handle: int) -> None:
        """StepLauncher callback (once per step, first bucket): upload the hyper-parameters on the raw stream ``handle``."""
        if self._comm_stream is not None and handle == self._comm_stream.cuda_stream:
            st = self._comm_stream
        else:
            st = torch.cuda.ExternalStream(handle, device=self.device)
        self._upload_hyper(st)

    def _launch_bucket(self, b: int) -> None:
        cur = torch.cuda.current_stream(self.device)
        stream = self._comm_stream if self._comm_stream is not None else cur
        if self._comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(cur)
            stream.wait_event(ev)
        from ..ops.functional import wgrad_join

        wgrad_join(stream)          # weight gradients are produced on a side stream (ops.functional.run_wgrad)
        if not self._hyper_uploaded:
            self._upload_hyper(stream)
            self._hyper_uploaded = True
        start, numel = int(self.plan["bucket_start"][b]), int(self.plan["bucket_numel"][b])
        blocks = max(1, min(self.comm_blocks, (numel // self.world + 2047) // 2048))
        if self.world == 1:
            self.C.fused_sgd_local(self.W.data_ptr() + start * 4, self.G.data_ptr() + start * 4,
                                   self.M.data_ptr() + start * 4, self.Wb.data_ptr() + start * 2,
                                   self._hyper_dev.data_ptr(), numel, max(blocks, min(4 * self._sms, numel // 2048 + 1)),
                                   stream.cuda_stream)
        else:
            tail = self._scalars_pending > 0 and b == self.num_buckets - 1
            oneshot = numel * 4 <= self.oneshot_bytes
            if oneshot:
                blocks = max(1, min(self.comm_blocks, (numel + 2047) // 2048))
            self.C.fused_allreduce_sgd(self.ctx, start, numel, self.M.data_ptr(), self._hyper_dev.data_ptr(), 0,
                                       self.use_mc, self.wire_bf16, blocks, stream.cuda_stream,
                                       self.arena.offsets["scalars"] if tail else 0,
                                       self._scalars_out.data_ptr() if tail else 0, oneshot,
                                       b == self.num_buckets - 1)

    def _on_ready(self, idx: int) -> None:
        if self._ready_seen[idx]:
            return
        self._ready_seen[idx] = True
        b = int(self.plan["param_bucket"][idx])
        self._pending[b] -= 1
        if self._hold_buckets:          # selfcheck: keep the gradients in the arena until step() (no overlap)
            return
        # buckets are launched strictly in plan order so every rank issues the same kernel sequence
        while self._next_bucket < self.num_buckets and self._pending[self._next_bucket] == 0:
            self._launch_bucket(self._next_bucket)
            self._next_bucket += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # parameters that received no gradient this step still take part (their slots are zero)
        if self._launcher is not None:
            self._launcher.finish()                      # ... and joins the communication stream into the current one
            _ext.add_launches(self._launcher.launches - self._launches_seen)
            self._launches_seen = self._launcher.launches
        else:
            while self._next_bucket < self.num_buckets:
                self._launch_bucket(self._next_bucket)
                self._next_bucket += 1
            if self._comm_stream is not None:
                ev = torch.cuda.Event()
                ev.record(self._comm_stream)
                torch.cuda.current_stream(self.device).wait_event(ev)
        if self._scalars_pending and self.world == 1:
            self._scalars_out.copy_(self._scalars_in)
        self._scalars_last, self._scalars_pending = self._scalars_pending, 0
        from ..ops import fp8 as _fp8

        _fp8.end_of_step()          # fp8 mode: this step's amax values become next step's quantisation scales
        self._pending = list(self.plan["bucket_param_count"])
        self._ready_seen = [False] * len(self.params)
        self._next_bucket = 0
        self._hyper_uploaded = False
        if self._launcher is not None:
            self._launcher.reset()
        self._first_step = False
        self._steps += 1
        if self.debug:
            self._debug_check()
        return loss

    def _debug_check(self) -> None:
        """Debug build of the step (enable with ``debug=True`` / ``DDL_COMM_DEBUG=1``): joins the device, then checks
        that (1) no barrier timed out, (2) every gradient accumulator was cleared by its bucket kernel, (3) all blocks
        of the bucket channel agree on the flag epoch, (4) replicas of the weights are bit-identical across ranks;
        finally the bf16 wire staging is poisoned with NaNs so a stale read in the next step cannot go unnoticed."""
        torch.cuda.synchronize(self.device)
        self.check_errors()
        if float(self.G.abs().max()) != 0.0:
            raise RuntimeError(f"rank {self.rank}: gradient accumulators not cleared after step {self._steps}")
        if self.world > 1:
            # flag epochs: block 0 of the bucket channel took part in every launch on every rank -> same count everywhere
            e0 = float(self._epochs[0])
            if dist.allreduce_scalar(e0, op="max") != dist.allreduce_scalar(e0, op="min"):
                raise RuntimeError(f"flag epochs differ across ranks after step {self._steps} (rank {self.rank}: {e0})")
            for probe in (float(self.W.double().sum()), float(self.W.double().abs().sum())):
                if dist.allreduce_scalar(probe, op="max") != dist.allreduce_scalar(probe, op="min"):
                    raise RuntimeError(f"weight replicas diverged after step {self._steps} (rank {self.rank})")
        if self.wire_bf16:
            self.arena.region("stage", torch.bfloat16, self.total_elems).fill_(float("nan"))

    # ---- scalar piggy-back (reference ``PyTorch_hvd/src/imagenet_pytorch_horovod.py:246``) ----------
    @torch.no_grad()
    def piggyback(self, values: torch.Tensor) -> None:
        """Queue up to SCALAR_SLOTS device scalars (loss, accuracy, ...) to be averaged across ranks by the LAST bucket
        kernel of this step — the reference's two blocking per-step MPI allreduces ride along for free.  Call after
        the forward pass and before ``step()``; read the result with :meth:`averaged_scalars` after ``step()``."""
        v = values.detach().reshape(-1).to(torch.float32)
        if v.numel() > self._scalars_in.numel():
            raise ValueError(f"piggyback carries at most {self._scalars_in.numel()} scalars")
        self._scalars_in[:v.numel()].copy_(v)
        self._scalars_pending = int(v.numel())
        if self._launcher is not None:
            self._launcher.set_scalars_pending(True)

    def averaged_scalars(self) -> torch.Tensor:
        """Cross-rank means of the values passed to :meth:`piggyback` before the last ``step()`` (device tensor)."""
        return self._scalars_out[:self._scalars_last]

    def zero_grad(self, set_to_none: bool = False):
        """No-op: the fused kernel clears each gradient slot after consuming it."""
        return None

    def synchronize(self) -> None:
        torch.cuda.synchronize(self.device)
        self.check_errors()

    def check_errors(self) -> None:
        code = int(self._error.item())
        if code:
            raise RuntimeError(f"fused allreduce barrier timed out waiting for rank {code - 1} "
                               f"(rank {self.rank}); a peer died or diverged")

    # ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def refresh_bf16(self) -> None:
        """Recompute the bf16 compute copy from the fp32 masters (after load / broadcast)."""
        self.C.cast_f32_bf16(self.W.data_ptr(), self.Wb.data_ptr(), self.total_elems,
                             torch.cuda.current_stream(self.device).cuda_stream)

    @torch.no_grad()
    def broadcast_parameters(self, root_rank: int = 0) -> None:
        """K16: one broadcast kernel over the weight region (multimem.st / peer stores)."""
        if self.world == 1:
            return
        st = torch.cuda.current_stream(self.device)
        self.C.broadcast(self.ctx, 1, self.arena.offsets["weight"], self.total_elems * 4, root_rank, self.use_mc, 32,
                         st.cuda_stream)
        self.refresh_bf16()

    def broadcast_state(self, root_rank: int = 0) -> None:
        """``hvd.broadcast_optimizer_state`` equivalent (momentum + hyper-parameters)."""
        if self.world == 1:
            return
        groups = dist.broadcast_object([{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                                       root_rank)
        for g, src in zip(self.param_groups, groups):
            g.update(src)
        self._first_step = bool(dist.broadcast_object(self._first_step, root_rank))
        dist.broadcast_(self.M, root_rank)

    @torch.no_grad()
    def full_momentum(self) -> torch.Tensor:
        """Assemble the sharded momentum (rank r holds slice r of every bucket) on every rank."""
        if self.world == 1:
            return self.M
        scratch = self.G  # gradients are zero between steps; reuse the symmetric region as transport
        st = torch.cuda.current_stream(self.device)
        for b in range(self.num_buckets):
            start, numel = int(self.plan["bucket_start"][b]), int(self.plan["bucket_numel"][b])
            self.C.allgather_slices(self.ctx, 2, self.M.data_ptr(), self.arena.offsets["grad"], start, numel,
                                    self.use_mc, 16, st.cuda_stream)
        out = scratch.clone()
        self.C.barrier(self.ctx, 3, st.cuda_stream)
        scratch.zero_()
        self.M.copy_(out)
        return self.M

    def state_dict(self):
        """``torch.optim.SGD.state_dict()`` layout: ``state`` keyed by REGISTRATION-order parameter index,
        ``param_groups[0]['params'] = [0..N-1]``, momentum under ``momentum_buffer`` — so checkpoints are
        interchangeable with the reference's ``optimizer.state_dict()``
        (``PyTorch_hvd/src/imagenet_pytorch_horovod.py:228-235``) and with this repo's CPU optimizer.  The engine's
        own bookkeeping rides in an extra ``ddl`` key that ``torch.optim.SGD.load_state_dict`` ignores."""
        mom = self.full_momentum()
        n = len(self.params)
        state = {}
        g0 = self.param_groups[0]
        if float(g0["momentum"]) != 0.0 and not self._first_step:
            for i, p in enumerate(self.params):               # self.params is reverse registration order
                o, cnt = int(self.plan["param_offset"][i]), p.numel()
                state[n - 1 - i] = {"momentum_buffer": _param_view(mom[o:o + cnt], p).detach().clone().cpu()}
        groups = []
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            for k, v in (("maximize", False), ("foreach", None), ("differentiable", False), ("fused", None)):
                d.setdefault(k, v)                 # keys torch.optim.SGD.step() expects after load_state_dict
            d["params"] = list(range(n))
            groups.append(d)
        return {"state": state, "param_groups": groups, "ddl": {"fused": True, "first_step": self._first_step}}

    def load_state_dict(self, sd):
        """Accepts the torch layout written by :meth:`state_dict` / ``torch.optim.SGD`` (registration order) and the
        round-1 private layout (``order == 'reverse_registration'``); shapes are validated before anything is copied."""
        n = len(self.params)
        legacy = sd.get("order") == "reverse_registration"
        st = sd.get("state", {})

        def entry(reg_index: int):
            key = (n - 1 - reg_index) if legacy else reg_index
            return st.get(key, st.get(str(key)))

        plan = []
        for i, p in enumerate(self.params):
            ent = entry(n - 1 - i)
            buf = None if ent is None else ent.get("momentum_buffer")
            if buf is None:
                continue
            if tuple(buf.shape) != tuple(p.shape):
                raise ValueError(f"optimizer state for parameter {n - 1 - i} has shape {tuple(buf.shape)}, "
                                 f"expected {tuple(p.shape)} (checkpoint from a different model or parameter order)")
            plan.append((i, buf))
        for g, src in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in src.items() if k != "params"})
        with torch.no_grad():
            for i, buf in plan:
                p = self.params[i]
                o, cnt = int(self.plan["param_offset"][i]), p.numel()
                _param_view(self.M[o:o + cnt], p).copy_(buf.to(self.device))
        meta = sd.get("ddl", {})
        if "first_step" in meta:
            self._first_step = bool(meta["first_step"])
        elif "first_step" in sd:
            self._first_step = bool(sd["first_step"])
        else:
            self._first_step = len(plan) == 0 and float(self.param_groups[0]["momentum"]) != 0.0
        self._hyper_blob = None
        self.refresh_bf16()

    # ---------------------------------------------------------------------------------------------
    def describe(self) -> str:
        mb = self.total_elems * 4 / (1 << 20)
        one = sum(1 for n in self.plan["bucket_numel"] if int(n) * 4 <= self.oneshot_bytes) if self.world > 1 else 0
        return (f"FusedSGD(world={self.world}, params={len(self.params)}, buckets={self.num_buckets}"
                f"{f' ({one} one-shot)' if one else ''}, "
                f"arena={mb:.1f} MiB fp32, wire={'bf16' if self.wire_bf16 else 'fp32'}, "
                f"transport={'nvls-multicast' if self.use_mc else ('p2p' if self.world > 1 else 'local')})")
