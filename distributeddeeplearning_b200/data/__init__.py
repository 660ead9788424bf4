"""Input pipelines: synthetic (host FakeData + device Philox), ImageFolder, samplers, ImageNet prep."""
from .synthetic import DeviceSyntheticLoader, FakeData, fixed_synthetic_batch  # noqa: F401
from .sampler import DistributedSampler, get_sampler  # noqa: F401
