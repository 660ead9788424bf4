"""Minimal python-fire replacement: map ``--key value`` / ``--key=value`` argv onto a function's kwargs.

Parity: the reference exposes ``main`` through ``fire.Fire(main)``
(``PyTorch_imagenet/src/imagenet_pytorch_horovod.py:444-446``; ``control/src/aml_compute.py:638-647``);
AzureML turns ``script_params`` into ``--key value`` argv.  python-fire is not in this image.
Values are coerced like fire does: ints, floats, True/False/None literals, else strings; keys accept
both ``--warmup_epochs`` and ``--warmup-epochs``; bare ``--flag`` means True.
"""
from __future__ import annotations

import ast
import inspect
import sys
from typing import Any, Callable, Dict, List, Optional


def _coerce(v: str) -> Any:
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        low = v.lower()
        if low in ("true", "false"):
            return low == "true"
        if low == "none":
            return None
        return v


def parse_kwargs(fn: Callable, argv: List[str]) -> Dict[str, Any]:
    params = inspect.signature(fn).parameters
    out: Dict[str, Any] = {}
    i = 0
    while i < len(argv):
        tok = argv[i]
        if not tok.startswith("--"):
            raise SystemExit(f"unexpected positional argument {tok!r}")
        key, eq, val = tok[2:].partition("=")
        key = key.replace("-", "_")
        if key in ("help", "h"):
            print(f"usage: {fn.__name__} " + " ".join(f"[--{k} {p.default!r}]" for k, p in params.items()))
            raise SystemExit(0)
        if key not in params:
            raise SystemExit(f"unknown option --{key}; valid: {', '.join(params)}")
        if eq:
            out[key] = _coerce(val)
            i += 1
        elif i + 1 < len(argv) and not argv[i + 1].startswith("--"):
            out[key] = _coerce(argv[i + 1])
            i += 2
        else:
            out[key] = True
            i += 1
    return out


def Fire(fn: Callable, argv: Optional[List[str]] = None):
    return fn(**parse_kwargs(fn, list(sys.argv[1:] if argv is None else argv)))
