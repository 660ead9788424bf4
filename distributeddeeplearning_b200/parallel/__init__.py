"""Data-parallel runtime: Horovod-equivalent API (dist), compression, optimizers, fused engine."""
from . import dist  # noqa: F401
from .compression import Compression  # noqa: F401
from .dist import (  # noqa: F401
    allreduce,
    barrier,
    broadcast,
    broadcast_object,
    broadcast_optimizer_state,
    broadcast_parameters,
    init,
    local_rank,
    rank,
    shutdown,
    size,
)
from .optimizer import DistributedOptimizer, HookedDistributedOptimizer  # noqa: F401
