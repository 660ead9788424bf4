"""Logging configuration.

Parity: the reference selects an ini file through the ``LOG_CONFIG`` env var and feeds it to
``logging.config.fileConfig`` (``PyTorch_imagenet/src/imagenet_pytorch_horovod.py:445``;
ini files ``control/src/logging.conf``, ``PyTorch_imagenet/src/logging.conf`` (INFO),
``PyTorch_hvd/src/logging.conf`` (DEBUG)).  Same contract here; without ``LOG_CONFIG`` a
built-in stdout config is used.  Packaged ini twins live in ``utils/logging_*.conf``.
"""
from __future__ import annotations

import logging
import logging.config
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
PACKAGED = {
    "control": os.path.join(_HERE, "logging_control.conf"),
    "imagenet": os.path.join(_HERE, "logging_imagenet.conf"),
    "hvd": os.path.join(_HERE, "logging_hvd.conf"),
}


def configure(default: str = "imagenet", level: int | None = None) -> str:
    """Configure logging; returns the path used ('' for the built-in fallback)."""
    path = os.getenv("LOG_CONFIG", "") or PACKAGED.get(default, "")
    if path and os.path.isfile(path):
        logging.config.fileConfig(path, disable_existing_loggers=False)
    else:
        path = ""
        logging.basicConfig(stream=sys.stdout, level=logging.INFO,
                            format="%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    if level is not None:
        logging.getLogger().setLevel(level)
    return path
