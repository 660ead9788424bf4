"""Host and device timers.

Parity: the reference ships one host wall-clock ``Timer`` (context manager / start /
stop / ``elapsed``) plus a ``@timer`` decorator, copied three times
(reference ``PyTorch_imagenet/src/timer.py:7-105``).  Here the same surface exists once,
and a ``DeviceTimer`` (CUDA events on the launching stream, max over ranks) is added
because every number this repo reports must be device-timed.
"""
from __future__ import annotations

import functools
import time
from typing import Callable, List, Optional

_FMT = "{prefix}took {elapsed:.{round}f} seconds"


class TimerError(RuntimeError):
    pass


class Timer:
    """Wall-clock timer usable as context manager or via start()/stop()."""

    def __init__(self, output: Optional[Callable[[str], None]] = None, fmt: str = _FMT,
                 prefix: str = "", round_to: int = 3):
        self._output = output
        self._fmt = fmt
        self._prefix = prefix
        self._round = round_to
        self._t0: Optional[float] = None
        self._t1: Optional[float] = None
        self.running = False

    def start(self) -> "Timer":
        self._t0 = time.perf_counter()
        self._t1 = None
        self.running = True
        return self

    def stop(self) -> "Timer":
        if self._t0 is None:
            raise TimerError("stop() called before start()")
        self._t1 = time.perf_counter()
        self.running = False
        return self

    @property
    def elapsed(self) -> float:
        if self._t0 is None:
            raise TimerError("timer never started")
        end = self._t1 if self._t1 is not None else time.perf_counter()
        return end - self._t0

    def __enter__(self) -> "Timer":
        return self.start()

    def __exit__(self, *exc) -> bool:
        self.stop()
        if self._output is not None:
            self._output(str(self))
        return False

    def __str__(self) -> str:
        return self._fmt.format(prefix=self._prefix, elapsed=self.elapsed, round=self._round)


def timer(output: Optional[Callable[[str], None]] = print, fmt: str = _FMT, prefix: str = "",
          round_to: int = 3):
    """Decorator: time every call of the wrapped function and report through ``output``."""

    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **kw):
            with Timer(output=output, fmt=fmt, prefix=prefix or f"{fn.__name__} ", round_to=round_to):
                return fn(*a, **kw)

        return wrapped

    return deco


class DeviceTimer:
    """CUDA-event timer on the current stream; host perf_counter on CPU.

    ``elapsed_ms`` synchronises only on the stop event.  ``max_over_ranks`` reduces with the
    process group so multi-GPU numbers are the slowest rank's, never wall clock.
    """

    def __init__(self, device=None):
        import torch

        self._torch = torch
        self._cuda = torch.cuda.is_available() and (device is None or str(device).startswith("cuda"))
        self._spans: List[tuple] = []
        self._open = None

    def start(self):
        if self._cuda:
            ev = self._torch.cuda.Event(enable_timing=True)
            ev.record()
            self._open = ev
        else:
            self._open = time.perf_counter()
        return self

    def stop(self):
        if self._open is None:
            raise TimerError("stop() before start()")
        if self._cuda:
            ev = self._torch.cuda.Event(enable_timing=True)
            ev.record()
            self._spans.append((self._open, ev))
        else:
            self._spans.append((self._open, time.perf_counter()))
        self._open = None
        return self

    def elapsed_ms(self) -> float:
        total = 0.0
        for a, b in self._spans:
            if self._cuda:
                b.synchronize()
                total += a.elapsed_time(b)
            else:
                total += (b - a) * 1e3
        return total

    def reset(self):
        self._spans.clear()
        self._open = None

    def max_over_ranks(self) -> float:
        from ..parallel import dist

        return dist.allreduce_scalar(self.elapsed_ms(), op="max")
