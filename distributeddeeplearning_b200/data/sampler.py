"""Per-rank data sharding.

Parity: ``torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=hvd.size(),
rank=hvd.rank())`` when distributed, ``RandomSampler`` otherwise
(``imagenet_pytorch_horovod.py:248-254``; ``PyTorch_hvd/...:94-95,108-109``); ``set_epoch``
reseeds the shuffle.  Each rank sees ``ceil(len / size)`` samples (tail padded by wrap-around).
"""
from __future__ import annotations

import math
from typing import Iterator

import torch
from torch.utils.data import RandomSampler, Sampler

from ..parallel import dist


class DistributedSampler(Sampler):
    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=True, seed=0):
        self.dataset = dataset
        self.num_replicas = dist.size() if num_replicas is None else num_replicas
        self.rank = dist.rank() if rank is None else rank
        self.shuffle, self.seed, self.epoch = shuffle, seed, 0
        self.num_samples = math.ceil(len(dataset) / self.num_replicas)
        self.total_size = self.num_samples * self.num_replicas

    def __iter__(self) -> Iterator[int]:
        n = len(self.dataset)
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            idx = torch.randperm(n, generator=g).tolist()
        else:
            idx = list(range(n))
        pad = self.total_size - len(idx)
        if pad > 0:
            idx += (idx * math.ceil(pad / max(len(idx), 1)))[:pad]
        return iter(idx[self.rank:self.total_size:self.num_replicas])

    def __len__(self) -> int:
        return self.num_samples

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch


def get_sampler(dataset, is_distributed=None, shuffle=True):
    is_distributed = dist.is_distributed() if is_distributed is None else is_distributed
    if is_distributed:
        return DistributedSampler(dataset, shuffle=shuffle)
    return RandomSampler(dataset) if shuffle else None
