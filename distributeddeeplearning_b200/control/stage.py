"""Stage the framework's sources into a docker build context (``make build`` of a generated project).

    python -m distributeddeeplearning_b200.control.stage control/Docker/framework

Copies the package (python + csrc, no built ``.so``, no caches) so that the image recipe can compile the native module
inside the image (control/docker/dockerfile).  Reference: the control image of the reference installs its CLI
dependencies with conda/pip at build time (``control/Docker/dockerfile``); ours additionally has native code to build.
"""
from __future__ import annotations

import os
import shutil
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage(dest: str) -> str:
    dest = os.path.abspath(dest)
    out = os.path.join(dest, os.path.basename(_PKG))
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(dest, exist_ok=True)
    shutil.copytree(_PKG, out, ignore=shutil.ignore_patterns("*.so", "*.o", "__pycache__", "*.pyc", "build"))
    return out


if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    print("staged", stage(sys.argv[1]))
