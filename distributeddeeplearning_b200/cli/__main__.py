"""``python -m distributeddeeplearning_b200.cli <task> ...`` — same task grammar as ``inv``."""
import sys

from invoke import Program

from .tasks import namespace

program = Program(namespace=namespace, name="b200-ddl", version="0.1.0")

if __name__ == "__main__":
    sys.exit(program.run())
