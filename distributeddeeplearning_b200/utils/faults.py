"""Fault injection + failure-detection helpers (the reference has none).

``DDL_INJECT_FAULT="rank:step[:kind]"`` makes that rank fail at that training step
(kind = ``exit`` (default, exit code 17), ``raise``, ``hang``).  Used by tests to prove the launcher
tears the whole job down on the first dead rank instead of hanging in a collective, and that the
fused allreduce barrier's bounded spin raises on the survivors.
"""
from __future__ import annotations

import os
import sys
import time


def maybe_inject(step: int, rank: int) -> None:
    spec = os.environ.get("DDL_INJECT_FAULT")
    if not spec:
        return
    parts = spec.split(":")
    try:
        frank, fstep = int(parts[0]), int(parts[1])
    except (ValueError, IndexError):
        return
    kind = parts[2] if len(parts) > 2 else "exit"
    if rank != frank or step != fstep:
        return
    sys.stderr.write(f"[fault-injection] rank {rank} failing at step {step} ({kind})\n")
    sys.stderr.flush()
    if kind == "raise":
        raise RuntimeError(f"injected fault on rank {rank} at step {step}")
    if kind == "hang":
        time.sleep(10 ** 6)
    os._exit(17)
