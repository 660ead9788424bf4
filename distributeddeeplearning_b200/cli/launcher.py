"""Local job launcher + rank supervisor (replaces the AzureML/MPI launch path).

Parity: the reference submits ``entry_script`` + ``script_params`` to an AzureML Estimator with
``distributed_backend="mpi"``, ``node_count`` x ``process_count_per_node`` ranks, NCCL/``DISTRIBUTED``
env vars and streams the rank logs back (``control/src/aml_compute.py:74-133,495-526``); "local"
mode runs exactly one non-distributed process (``:434-445``).  Here:

* ``launch(module, argv, gpus=N)`` starts one Python process per GPU of this box with
  ``RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT`` (env rendezvous, no MPI) plus the
  reference's ``DISTRIBUTED`` flag, inside a fresh process group;
* children's stdout/stderr are streamed line by line, prefixed ``[rank k]`` (rank 0 unprefixed);
* the supervisor polls the children; on the first non-zero exit it terminates the whole group,
  prints the failing rank's tail of stderr and returns its exit code — a dead
  rank never leaves the others hanging in a collective;
* an optional wall-clock ``timeout`` bounds the job; ``--inject-fault rank:step`` style testing is
  supported by exporting ``DDL_INJECT_FAULT`` to the children (see ``utils.faults``);
* every launch is recorded as a run directory (``utils.runs.Run``): argv, world size, status, exit code.
"""
from __future__ import annotations

import os
import signal
import socket
import subprocess
import sys
import threading
import time
from collections import deque
from typing import Dict, List, Optional, Sequence

from ..utils.runs import Run


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Pump(threading.Thread):
    def __init__(self, stream, prefix: str, sink, keep: int = 60):
        super().__init__(daemon=True)
        self.stream, self.prefix, self.sink = stream, prefix, sink
        self.tail: deque = deque(maxlen=keep)

    def run(self):
        for raw in iter(self.stream.readline, b""):
            line = raw.decode("utf-8", "replace").rstrip("\n")
            self.tail.append(line)
            try:
                self.sink.write(f"{self.prefix}{line}\n")
                self.sink.flush()
            except ValueError:
                break
        self.stream.close()


class LaunchResult:
    def __init__(self, returncode: int, failed_rank: Optional[int], run: Optional[Run], elapsed: float):
        self.returncode, self.failed_rank, self.run, self.elapsed = returncode, failed_rank, run, elapsed

    def __repr__(self):
        return f"LaunchResult(returncode={self.returncode}, failed_rank={self.failed_rank}, elapsed={self.elapsed:.1f}s)"


def launch(module: str, argv: Sequence[str] = (), gpus: int = 1, no_cuda: bool = False,
           env: Optional[Dict[str, str]] = None, experiment: Optional[str] = None, timeout: Optional[float] = None,
           master_port: Optional[int] = None, stdout=None, stderr=None, record: bool = True,
           python: Optional[str] = None, poll_s: float = 0.1, grace_s: float = 5.0) -> LaunchResult:
    """Run ``python -m <module> <argv>`` on ``gpus`` ranks of this box and supervise them."""
    stdout = stdout or sys.stdout
    stderr = stderr or sys.stderr
    world = max(1, int(gpus))
    port = master_port or free_port()
    run = None
    if record:
        run = Run(experiment or module.rsplit(".", 1)[-1])
        run.set(argv=list(argv), module=module, world_size=world, no_cuda=no_cuda)
    base_env = dict(os.environ)
    base_env.update(env or {})
    base_env.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "WORLD_SIZE": str(world),
                     "DISTRIBUTED": "True" if world > 1 else "False", "PYTHONUNBUFFERED": "1"})
    if no_cuda:
        base_env["DDL_NO_CUDA"] = "1"
    if run is not None:
        base_env["DDL_RUN_DIR"] = run.dir
    procs: List[subprocess.Popen] = []
    pumps: List[_Pump] = []
    t0 = time.time()
    try:
        for r in range(world):
            e = dict(base_env)
            e.update({"RANK": str(r), "LOCAL_RANK": str(r)})
            cmd = [python or sys.executable, "-m", module, *argv]
            p = subprocess.Popen(cmd, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, start_new_session=True)
            procs.append(p)
            pre = "" if r == 0 else f"[rank {r}] "
            po, pe = _Pump(p.stdout, pre, stdout), _Pump(p.stderr, pre, stderr)
            po.start()
            pe.start()
            pumps += [po, pe]
        failed_rank, code = None, 0
        while True:
            alive = False
            for r, p in enumerate(procs):
                rc = p.poll()
                if rc is None:
                    alive = True
                elif rc != 0 and failed_rank is None:
                    failed_rank, code = r, rc
            if failed_rank is not None or not alive:
                break
            if timeout is not None and time.time() - t0 > timeout:
                failed_rank, code = -1, 124
                break
            time.sleep(poll_s)
        if failed_rank is not None:
            _terminate(procs, grace_s)
            if failed_rank >= 0:
                tail = list(pumps[2 * failed_rank + 1].tail)[-20:]
                stderr.write(f"[launcher] rank {failed_rank} exited with code {code}; terminated the job. "
                             f"Last stderr lines of that rank:\n" + "\n".join("    " + t for t in tail) + "\n")
            else:
                stderr.write(f"[launcher] job exceeded the {timeout:.0f}s limit; terminated.\n")
        for pm in pumps:
            pm.join(timeout=2.0)
        if run is not None:
            run.set(exit_code=code, failed_rank=failed_rank, elapsed_s=round(time.time() - t0, 3))
            run.complete("Completed" if code == 0 else "Failed")
        return LaunchResult(code, failed_rank, run, time.time() - t0)
    except BaseException:
        _terminate(procs, grace_s)
        if run is not None:
            run.fail("launcher interrupted")
        raise


def _terminate(procs: List[subprocess.Popen], grace_s: float) -> None:
    """SIGTERM each child's own process group (exact pgids we created), then SIGKILL stragglers."""
    for p in procs:
        if p.poll() is None:
            try:
                os.killpg(p.pid, signal.SIGTERM)
            except (ProcessLookupError, PermissionError):
                pass
    deadline = time.time() + grace_s
    for p in procs:
        while p.poll() is None and time.time() < deadline:
            time.sleep(0.05)
        if p.poll() is None:
            try:
                os.killpg(p.pid, signal.SIGKILL)
            except (ProcessLookupError, PermissionError):
                pass
            p.wait()
