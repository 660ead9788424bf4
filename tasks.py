"""`inv --list` entry point (the reference ships a root tasks.py; ours just re-exports the tree)."""
from distributeddeeplearning_b200.cli.tasks import namespace  # noqa: F401

ns = namespace
