"""Local run history + TensorBoard scalars.

Parity: the reference logs to AzureML run history (``Run.get_context()``, ``run.tag``,
``run.log_row`` — ``PyTorch_imagenet/src/imagenet_pytorch_horovod.py:321-323,425,434``),
to TensorBoard via tensorboardX (``:329,426-436``; ``PyTorch_hvd/...:78,172-174``), and
offers ``inv runs`` / ``inv experiments`` / ``inv tensorboard`` viewers (``tasks.py:120-168``).
Here a run is a directory ``<RUNS_DIR>/<experiment>/<run_id>/`` holding ``run.json``
(tags, status, argv), ``metrics.jsonl`` (one row per ``log_row``), per-rank JSONL records
(device-timed) and ``tb/`` event files; the viewers list / tail those directories.
"""
from __future__ import annotations

import json
import os
import time
import uuid
from typing import Any, Dict, List, Optional


def _now() -> str:
    return time.strftime("%Y-%m-%dT%H:%M:%S")


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass


def summary_writer(logdir: str):
    """TensorBoard writer (torch.utils.tensorboard); a no-op object if unavailable."""
    try:
        from torch.utils.tensorboard import SummaryWriter

        return SummaryWriter(log_dir=logdir)
    except Exception:  # pragma: no cover - tensorboard not installed
        return _NullWriter()


class Run:
    """One experiment run (the local stand-in for ``azureml.core.run.Run``)."""

    def __init__(self, experiment: str, root: Optional[str] = None, run_id: Optional[str] = None,
                 create: bool = True):
        root = root or os.getenv("DDL_RUNS_DIR") or os.getenv("RUNS_DIR") or "runs"
        self.experiment = experiment
        self.id = run_id or (time.strftime("%Y%m%d-%H%M%S") + "-" + uuid.uuid4().hex[:6])
        self.dir = os.path.join(root, experiment, self.id)
        self._meta: Dict[str, Any] = {"experiment": experiment, "id": self.id, "status": "Running",
                                      "started": _now(), "tags": {}}
        if create:
            os.makedirs(self.dir, exist_ok=True)
            self._flush()

    # ---- AzureML-like surface -------------------------------------------------------
    @classmethod
    def get_context(cls, experiment: Optional[str] = None) -> "Run":
        """Run chosen by the launcher through DDL_RUN_DIR / DDL_EXPERIMENT, else an ad-hoc one."""
        d = os.getenv("DDL_RUN_DIR")
        if d:
            exp = os.path.basename(os.path.dirname(d.rstrip("/")))
            run = cls(exp, root=os.path.dirname(os.path.dirname(d.rstrip("/"))),
                      run_id=os.path.basename(d.rstrip("/")))
            return run
        return cls(experiment or os.getenv("DDL_EXPERIMENT", "adhoc"))

    def tag(self, key: str, value: Any = None):
        self._meta["tags"][key] = value
        self._flush()

    def log(self, name: str, value: Any):
        self.log_row(name, value=value)

    def log_row(self, name: str, **cols):
        with open(os.path.join(self.dir, "metrics.jsonl"), "a") as f:
            f.write(json.dumps({"table": name, "ts": _now(), **cols}) + "\n")

    def complete(self, status: str = "Completed"):
        self._meta["status"] = status
        self._meta["ended"] = _now()
        self._flush()

    def fail(self, why: str = ""):
        self._meta["error"] = why
        self.complete("Failed")

    # ---- helpers ---------------------------------------------------------------------
    def set(self, **kv):
        self._meta.update(kv)
        self._flush()

    def rank_record(self, rank: int, **cols):
        with open(os.path.join(self.dir, f"rank{rank}.jsonl"), "a") as f:
            f.write(json.dumps({"ts": _now(), **cols}) + "\n")

    def tensorboard_dir(self) -> str:
        return os.path.join(self.dir, "tb")

    def _flush(self):
        tmp = os.path.join(self.dir, "run.json.tmp")
        with open(tmp, "w") as f:
            json.dump(self._meta, f, indent=1, default=str)
        os.replace(tmp, os.path.join(self.dir, "run.json"))


def list_experiments(root: str = "runs") -> List[str]:
    if not os.path.isdir(root):
        return []
    return sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d)))


def list_runs(experiment: str, root: str = "runs", last: Optional[int] = None) -> List[Dict[str, Any]]:
    base = os.path.join(root, experiment)
    out = []
    if not os.path.isdir(base):
        return out
    for rid in sorted(os.listdir(base)):
        p = os.path.join(base, rid, "run.json")
        if os.path.isfile(p):
            try:
                with open(p) as f:
                    out.append(json.load(f))
            except (OSError, ValueError):
                out.append({"id": rid, "status": "Unknown"})
    return out[-last:] if last else out


def read_metrics(experiment: str, run_id: str, root: str = "runs") -> List[Dict[str, Any]]:
    p = os.path.join(root, experiment, run_id, "metrics.jsonl")
    if not os.path.isfile(p):
        return []
    with open(p) as f:
        return [json.loads(line) for line in f if line.strip()]
