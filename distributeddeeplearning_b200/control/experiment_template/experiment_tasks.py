"""Submit tasks for your experiment (reference: PyTorch_experiment/pytorch_experiment.py)."""
import os

from invoke import Collection, task

from distributeddeeplearning_b200.cli import launcher

_BASE = os.path.dirname(os.path.abspath(__file__))


@task
def submit_local(c):
    raise NotImplementedError("point this at your module, e.g. launcher.launch('my_pkg.train_model', [...], gpus=1)")


@task
def submit_remote(c, node_count=8):
    raise NotImplementedError("launcher.launch('my_pkg.train_model', [...], gpus=node_count)")


namespace = Collection("experiment", submit_local, submit_remote)
