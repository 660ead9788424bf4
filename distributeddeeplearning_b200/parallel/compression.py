"""Gradient wire-format compression.

Parity: ``hvd.Compression.none`` / ``hvd.Compression.fp16`` selected by ``--fp16-allreduce``
(reference ``PyTorch_benchmark/src/pytorch_synthetic_benchmark.py:19-23,69``;
``PyTorch_imagenet/...:300,398``; ``PyTorch_hvd/...:25-26,126``).  Horovod casts each
gradient fp32->fp16 before the allreduce and back after.  Here:

* ``none``  fp32 on the wire (exact parity with the reference default)
* ``fp16``  IEEE half on the wire (reference flag semantics)
* ``bf16``  bfloat16 on the wire — what the fused NVLink kernel reduces natively with
            ``multimem.ld_reduce...acc::f32...bf16x2`` (fp32 accumulate in the switch)

In the fused engine the cast(+1/N scale) is a prologue of the allreduce kernel
(SURVEY.md K13), not a separate per-tensor pass; these classes are the generic-path
implementation and the name registry.
"""
from __future__ import annotations

import torch


class Compressor:
    name = "none"
    wire_dtype = None  # None = keep tensor dtype

    @staticmethod
    def compress(tensor: torch.Tensor):
        return tensor, None

    @staticmethod
    def decompress(tensor: torch.Tensor, ctx):
        return tensor


class NoneCompressor(Compressor):
    pass


class _CastCompressor(Compressor):
    @classmethod
    def compress(cls, tensor: torch.Tensor):
        if tensor.dtype.is_floating_point and tensor.dtype != cls.wire_dtype:
            return tensor.to(cls.wire_dtype), tensor.dtype
        return tensor, None

    @staticmethod
    def decompress(tensor: torch.Tensor, ctx):
        return tensor.to(ctx) if ctx is not None else tensor


class FP16Compressor(_CastCompressor):
    name = "fp16"
    wire_dtype = torch.float16


class BF16Compressor(_CastCompressor):
    name = "bf16"
    wire_dtype = torch.bfloat16


class Compression:
    """Namespace mirroring ``hvd.Compression``."""

    none = NoneCompressor
    fp16 = FP16Compressor
    bf16 = BF16Compressor

    @staticmethod
    def by_name(name: str):
        try:
            return {"none": NoneCompressor, "fp16": FP16Compressor, "bf16": BF16Compressor}[name]
        except KeyError:
            raise ValueError(f"unknown compression {name!r} (none|fp16|bf16)") from None
