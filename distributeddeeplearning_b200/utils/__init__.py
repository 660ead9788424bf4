"""Utilities: timers, meters, LR schedule, config, logging, run history, checkpoints."""
from .meters import AverageMeter, Metric, accuracy, top1_accuracy  # noqa: F401
from .timer import DeviceTimer, Timer, timer  # noqa: F401
