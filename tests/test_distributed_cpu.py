"""Multi-process CPU/gloo plumbing (BASELINE config #1): launcher, Horovod-equivalent API, hooked DP optimizer,
checkpoint/resume, fault handling.  world_size=2, rendezvous on 127.0.0.1."""
import io
import os
import re
import subprocess
import sys
import textwrap

import pytest
import torch

from distributeddeeplearning_b200.cli import launcher

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_script(tmp_path, body, gpus=2, timeout=240, env=None):
    script = tmp_path / "w.py"
    script.write_text("import os, sys\nsys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(body))
    out, err = io.StringIO(), io.StringIO()
    e = {"PYTHONPATH": str(tmp_path) + os.pathsep + ROOT, "OMP_NUM_THREADS": "2"}
    e.update(env or {})
    res = launcher.launch("w", [], gpus=gpus, no_cuda=True, env=e, record=False, stdout=out, stderr=err, timeout=timeout)
    return res, out.getvalue(), err.getvalue()


def test_benchmark_golden_output_two_ranks():
    out, err = io.StringIO(), io.StringIO()
    res = launcher.launch("distributeddeeplearning_b200.workloads.benchmark",
                          ["--model", "resnet18", "--batch-size", "2", "--num-warmup-batches", "1",
                           "--num-batches-per-iter", "1", "--num-iters", "2", "--no-cuda"],
                          gpus=2, no_cuda=True, record=False, stdout=out, stderr=err, timeout=600,
                          env={"OMP_NUM_THREADS": "2"})
    assert res.returncode == 0, err.getvalue()
    lines = out.getvalue().strip().splitlines()
    assert lines[0] == "Model: resnet18"
    assert lines[1] == "Batch size: 2"
    assert lines[2] == "Number of CPUs: 2"
    assert lines[3] == "Running warmup..." and lines[4] == "Running benchmark..."
    assert re.fullmatch(r"Iter #0: \d+\.\d img/sec per CPU", lines[5])
    assert re.fullmatch(r"Iter #1: \d+\.\d img/sec per CPU", lines[6])
    assert re.fullmatch(r"Img/sec per CPU: \d+\.\d \+-\d+\.\d", lines[7])
    assert re.fullmatch(r"Total img/sec on 2 CPU\(s\): \d+\.\d \+-\d+\.\d", lines[8])
    assert len(lines) == 9            # rank 1 printed nothing


def test_collectives_and_hooked_optimizer(tmp_path):
    res, out, err = _run_script(tmp_path, """
        import torch
        from distributeddeeplearning_b200.parallel import dist, DistributedOptimizer, Compression
        dist.init()
        r, n = dist.rank(), dist.size()
        assert n == 2
        t = torch.full((4,), float(r + 1))
        assert dist.allreduce(t, average=True).tolist() == [1.5] * 4
        assert dist.allreduce(t, average=False).tolist() == [3.0] * 4
        assert dist.broadcast(torch.tensor([r + 10.0]), root_rank=1).item() == 11.0
        assert dist.broadcast_object({"rank": r}, 0) == {"rank": 0}
        assert dist.allreduce_scalar(float(r), op="max") == 1.0
        torch.manual_seed(r)                         # different init per rank on purpose
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
        for comp in (Compression.none, Compression.fp16):
            dopt = DistributedOptimizer(opt, named_parameters=model.named_parameters(), compression=comp)
            dist.broadcast_parameters(model.state_dict(), root_rank=0)
            dist.broadcast_optimizer_state(dopt, root_rank=0)
            w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
            assert torch.equal(dist.broadcast(w0, 0), w0)          # replicas identical after broadcast
            x = torch.randn(4, 8, generator=torch.Generator().manual_seed(100 + r))
            dopt.zero_grad()
            model(x).pow(2).mean().backward()
            local = [p.grad.clone() for p in model.parameters()]
            dopt.step()
            for p, g in zip(model.parameters(), local):           # grads were averaged across ranks
                avg = dist.allreduce(g, average=True)
                assert torch.allclose(p.grad, avg, atol=2e-3 if comp is Compression.fp16 else 1e-6)
            w1 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
            assert torch.allclose(dist.broadcast(w1, 0), w1, atol=1e-6)   # replicas still identical
        print("OK", r)
        dist.shutdown()
    """)
    assert res.returncode == 0, err
    assert "OK 0" in out and "OK 1" in out


def test_checkpoint_resume_roundtrip(tmp_path):
    res, out, err = _run_script(tmp_path, f"""
        import torch
        from distributeddeeplearning_b200.parallel import dist, DistributedOptimizer
        from distributeddeeplearning_b200.utils.checkpoint import save_checkpoint, load_checkpoint, agreed_resume_epoch
        dist.init()
        fmt = r"{tmp_path}/checkpoint-{{epoch}}.pth.tar"
        assert agreed_resume_epoch(fmt, 5) == 0
        torch.manual_seed(0)
        model = torch.nn.Linear(4, 3)
        opt = DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9))
        model(torch.ones(2, 4)).sum().backward(); opt.step()
        save_checkpoint(fmt.format(epoch=2), model, opt, epoch=2, step=7)
        dist.barrier()
        assert agreed_resume_epoch(fmt, 5) == 2
        torch.manual_seed(123 + dist.rank())
        model2 = torch.nn.Linear(4, 3)
        opt2 = DistributedOptimizer(torch.optim.SGD(model2.parameters(), lr=0.5, momentum=0.9))
        meta = load_checkpoint(fmt.format(epoch=2), model2, opt2)
        assert meta["epoch"] == 2 and meta["step"] == 7
        for a, b in zip(model.parameters(), model2.parameters()):
            assert torch.equal(a, b)
        assert opt2.param_groups[0]["lr"] == 0.1
        mb = [opt2.state[p]["momentum_buffer"] for p in model2.parameters()]
        ma = [opt.state[p]["momentum_buffer"] for p in model.parameters()]
        assert all(torch.equal(x, y) for x, y in zip(ma, mb))
        print("RESUMED", dist.rank())
        dist.shutdown()
    """)
    assert res.returncode == 0, err
    assert "RESUMED 0" in out and "RESUMED 1" in out


def test_launcher_tears_job_down_on_dead_rank(tmp_path):
    res, out, err = _run_script(tmp_path, """
        import time
        from distributeddeeplearning_b200.utils.faults import maybe_inject
        rank = int(os.environ["RANK"])
        for step in range(1000):
            maybe_inject(step, rank)
            time.sleep(0.05)
    """, env={"DDL_INJECT_FAULT": "1:3"}, timeout=120)
    assert res.returncode == 17 and res.failed_rank == 1
    assert "rank 1 exited with code 17" in err and "fault-injection" in err
    assert res.elapsed < 60                 # rank 0 (sleeping ~50 s) was terminated, not waited for


def test_launcher_timeout(tmp_path):
    res, out, err = _run_script(tmp_path, "import time; time.sleep(60)", gpus=1, timeout=2)
    assert res.returncode == 124 and "exceeded" in err


def test_imagenet_trainer_synthetic_cpu(tmp_path):
    out, err = io.StringIO(), io.StringIO()
    res = launcher.launch("distributeddeeplearning_b200.workloads.imagenet",
                          ["--epochs", "1", "--batch_size", "2", "--model", "resnet18", "--use_gpu", "False",
                           "--save_filepath", str(tmp_path / "ck-{epoch}.pt")],
                          gpus=2, no_cuda=True, record=False, stdout=out, stderr=err, timeout=900,
                          env={"FAKE_DATA_LENGTH": "8", "OMP_NUM_THREADS": "2", "PYTHONPATH": ROOT})
    assert res.returncode == 0, err.getvalue()[-3000:]
    text = out.getvalue()
    assert "Training epoch 0 took" in text and "Total images/sec:" in text and "setting lr to" in text
    assert "Distributed:      True" in text and "Num GPUs:         2.000" in text
    assert (tmp_path / "ck-1.pt").exists()


def test_hvd_trainer_resume_cpu(tmp_path):
    fmt = str(tmp_path / "checkpoint-{epoch}.pth.tar")
    common = ["--batch-size", "2", "--model", "resnet18", "--no-cuda", "--synthetic-length", "4",
              "--checkpoint-format", fmt, "--log-dir", str(tmp_path / "logs")]
    for epochs in ("1", "2"):
        out, err = io.StringIO(), io.StringIO()
        res = launcher.launch("distributeddeeplearning_b200.workloads.hvd_imagenet", ["--epochs", epochs] + common,
                              gpus=2, no_cuda=True, record=False, stdout=out, stderr=err, timeout=900,
                              env={"OMP_NUM_THREADS": "2", "PYTHONPATH": ROOT})
        assert res.returncode == 0, err.getvalue()[-3000:]
    assert os.path.exists(fmt.format(epoch=1)) and os.path.exists(fmt.format(epoch=2))
