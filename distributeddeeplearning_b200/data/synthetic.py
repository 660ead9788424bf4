"""Synthetic ImageNet-shaped data.

Parity:
* fixed random batch of the benchmark — ``torch.randn(B,3,224,224)`` + ``LongTensor(B).random_()%1000``
  moved once to the GPU (``pytorch_synthetic_benchmark.py:81-84``) -> ``fixed_synthetic_batch``: on
  CUDA the batch is generated ON the device by the Philox kernel directly as NHWC4 bf16 (K1).
* ``FakeData`` dataset — 640 distinct random images (32 x 20), ``length`` random indices into them,
  ``n_classes`` labels (``imagenet_pytorch_horovod.py:70-125``; size via ``FAKE_DATA_LENGTH``).
* ``DeviceSyntheticLoader`` — what the ImageNet trainer uses on CUDA in synthetic mode: same
  epoch length / per-rank sharding arithmetic as DataLoader+DistributedSampler, but every batch is
  produced by the device generator (no host pool, no per-step H2D).  ``host_pool=True`` reproduces
  the reference's host path (pinned pool + H2D per step) for end-to-end measurements.
"""
from __future__ import annotations

import math
import os
from typing import Iterator, Optional, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

_DATA_LENGTH = int(os.getenv("FAKE_DATA_LENGTH", 1281167))


def _create_data(batch_size, num_batches, dim, channels, seed=42):
    rng = np.random.RandomState(seed)
    return rng.rand(batch_size * num_batches, channels, dim[0], dim[1]).astype(np.float32)


def _create_labels(batch_size, num_batches, n_classes, seed=42):
    rng = np.random.RandomState(seed + 1)
    return rng.choice(n_classes, batch_size * num_batches)


class FakeData(Dataset):
    """Host-side fake dataset with the reference's shape: few distinct images, many indices."""

    def __init__(self, batch_size=32, num_batches=20, dim=(224, 224), n_channels=3, n_classes=10,
                 length: Optional[int] = None, data_transform=None, seed=42):
        self.dim, self.n_channels, self.n_classes = dim, n_channels, n_classes
        self._data = _create_data(batch_size, num_batches, dim, n_channels, seed)
        self._labels = _create_labels(batch_size, num_batches, n_classes, seed)
        self._length = int(length if length is not None else _DATA_LENGTH)
        self.translation_index = np.random.RandomState(seed + 2).choice(len(self._labels), self._length)
        self._transform = data_transform

    def __getitem__(self, idx):
        j = self.translation_index[idx]
        x = self._data[j]
        if self._transform is not None:
            x = self._transform(x)
        return x, int(self._labels[j])

    def __len__(self):
        return self._length


def fixed_synthetic_batch(batch_size: int, size: int = 224, classes: int = 1000, device=None, seed: int = 0):
    """One fixed (data, target) pair.  CUDA: Philox-generated NHWC4 bf16; CPU: torch.randn fp32 NCHW."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    if device.type == "cuda":
        from ..ops import native, use_native

        probe = torch.empty(0, device=device)
        if use_native(probe):
            data = native.philox_images(batch_size, size, size, seed, 0, device)
            target = native.philox_labels(batch_size, classes, seed + 1, 0, device)
            return data, target
    g = torch.Generator().manual_seed(seed)
    data = torch.randn(batch_size, 3, size, size, generator=g)
    target = torch.randint(0, classes, (batch_size,), generator=g)
    return data.to(device), target.to(device)


class DeviceSyntheticLoader:
    """Iterable of ``batches_per_epoch`` synthetic batches for THIS rank.

    ``len(loader)`` equals ``ceil(ceil(length / world) / batch)`` — what
    ``DataLoader(FakeData, sampler=DistributedSampler)`` yields per rank in the reference.
    """

    def __init__(self, length: int, batch_size: int, size: int = 224, classes: int = 1000, device=None,
                 rank: int = 0, world: int = 1, seed: int = 42, host_pool: bool = False, distinct: int = 640):
        self.length, self.batch_size, self.size, self.classes = int(length), batch_size, size, classes
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.rank, self.world, self.seed = rank, world, seed
        self.per_rank = math.ceil(self.length / world)
        self.epoch = 0
        self.host_pool = host_pool
        self._pool = None
        if host_pool:
            n = max(distinct, batch_size)
            g = torch.Generator().manual_seed(seed)
            self._pool = torch.rand(n, 3, size, size, generator=g)
            self._pool_labels = torch.randint(0, classes, (n,), generator=g)
            if self.device.type == "cuda":
                self._pool, self._pool_labels = self._pool.pin_memory(), self._pool_labels.pin_memory()

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __len__(self) -> int:
        return math.ceil(self.per_rank / self.batch_size)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        remaining = self.per_rank
        for i in range(len(self)):
            b = min(self.batch_size, remaining)
            remaining -= b
            if self._pool is not None:
                g = torch.Generator().manual_seed(self.seed + 7919 * self.epoch + 31 * i + self.rank)
                idx = torch.randint(0, self._pool.shape[0], (b,), generator=g)
                x, y = self._pool[idx], self._pool_labels[idx]
                if self.device.type == "cuda":
                    x, y = x.pin_memory().to(self.device, non_blocking=True), y.pin_memory().to(self.device, non_blocking=True)
                yield x, y
                continue
            step_seed = self.seed + 1000003 * self.epoch + 101 * i
            off = (self.rank * len(self) + i) * b * self.size * self.size
            if self.device.type == "cuda":
                from ..ops import native, use_native

                if use_native(torch.empty(0, device=self.device)):
                    yield (native.philox_images(b, self.size, self.size, step_seed, off, self.device),
                           native.philox_labels(b, self.classes, step_seed + 1, self.rank * 1_000_003 + i * b, self.device))
                    continue
            g = torch.Generator().manual_seed(step_seed + self.rank)
            yield (torch.rand(b, 3, self.size, self.size, generator=g).to(self.device),
                   torch.randint(0, self.classes, (b,), generator=g).to(self.device))
