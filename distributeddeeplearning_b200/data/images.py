"""Real-image (ImageFolder) pipeline.

Parity: train = RandomResizedCrop(224) -> RandomHorizontalFlip -> ToTensor -> Normalize,
val = Resize(256) -> CenterCrop(224) -> ToTensor -> Normalize, ``num_workers`` DataLoader workers with
pinned memory (``imagenet_pytorch_horovod.py:336-375``; ``PyTorch_hvd/...:82-111``).  Decode /
augment stays on CPU workers (not a named hot path); normalisation + NCHW->NHWC4 + bf16 cast run
in ONE device kernel (``ops.native.nchw_to_nhwc4``) when ``device_normalize`` is set, so the H2D
copy carries uint8-range floats only once and no fp32 normalised copy is materialised.
"""
from __future__ import annotations


import torch

RGB_MEAN = (0.485, 0.456, 0.406)
RGB_SD = (0.229, 0.224, 0.225)


def build_transforms(train: bool, size: int = 224, normalize_on_host: bool = True):
    from torchvision import transforms

    tail = [transforms.ToTensor()]
    if normalize_on_host:
        tail.append(transforms.Normalize(RGB_MEAN, RGB_SD))
    if train:
        return transforms.Compose([transforms.RandomResizedCrop(size), transforms.RandomHorizontalFlip(), *tail])
    return transforms.Compose([transforms.Resize(int(size * 256 / 224)), transforms.CenterCrop(size), *tail])


def image_folder_loader(path: str, batch_size: int, train: bool, size: int = 224, num_workers: int = 5,
                        device_normalize: bool = False, seed: int = 0):
    from torchvision import datasets

    from .sampler import get_sampler

    ds = datasets.ImageFolder(path, build_transforms(train, size, normalize_on_host=not device_normalize))
    sampler = get_sampler(ds, shuffle=train)
    loader = torch.utils.data.DataLoader(ds, batch_size=batch_size, sampler=sampler, num_workers=num_workers,
                                         pin_memory=torch.cuda.is_available(), drop_last=False,
                                         persistent_workers=num_workers > 0)
    return loader, sampler


class DeviceNormalizer:
    """(x - mean) / std + layout/dtype conversion in one device pass."""

    def __init__(self, device):
        self.mean = torch.tensor(RGB_MEAN, dtype=torch.float32, device=device)
        self.std = torch.tensor(RGB_SD, dtype=torch.float32, device=device)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        from ..ops import native, use_native

        if x.is_cuda and use_native(x):
            return native.nchw_to_nhwc4(x, self.mean, self.std)
        return (x - self.mean.view(1, 3, 1, 1).to(x.device)) / self.std.view(1, 3, 1, 1).to(x.device)
