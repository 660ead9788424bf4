import os
import time

import pytest
import torch

from distributeddeeplearning_b200.utils import config as cfg
from distributeddeeplearning_b200.utils import firelite, runs
from distributeddeeplearning_b200.utils.lr_schedule import adjust_learning_rate, learning_rate, lr_adjustment
from distributeddeeplearning_b200.utils.meters import AverageMeter, Metric, accuracy, top1_accuracy
from distributeddeeplearning_b200.utils.timer import DeviceTimer, Timer, TimerError, timer


def test_timer_context_and_decorator():
    lines = []
    with Timer(output=lines.append, prefix="epoch 0 ") as t:
        time.sleep(0.01)
    assert t.elapsed >= 0.01 and lines and lines[0].startswith("epoch 0 took")
    with pytest.raises(TimerError):
        Timer().stop()

    @timer(output=lines.append)
    def f(x):
        return x + 1

    assert f(1) == 2 and "f took" in lines[-1]


def test_device_timer_cpu():
    d = DeviceTimer(device="cpu").start()
    time.sleep(0.005)
    assert d.stop().elapsed_ms() >= 4.0


def test_lr_schedule_matches_reference_formula():
    # size=8, warmup 5 epochs, 100 batches/epoch: ramps base_lr -> base_lr*8, then step decay
    base, size, w, nb = 0.0125, 8, 5, 100
    assert learning_rate(base, 0, 0, nb, size, w) == pytest.approx(base * size * (1 / size) * ((1 / nb) * (size - 1) / w + 1))
    assert learning_rate(base, 4, nb - 1, nb, size, w) == pytest.approx(base * size)
    assert lr_adjustment(5, 0, nb, size, w) == 1.0
    assert lr_adjustment(30, 0, nb, size, w) == 1e-1
    assert lr_adjustment(60, 0, nb, size, w) == 1e-2
    assert lr_adjustment(80, 0, nb, size, w) == 1e-3
    # size 1: warm-up is flat
    assert learning_rate(base, 0, 0, nb, 1, w) == pytest.approx(base)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    assert adjust_learning_rate(opt, base, w, nb, 10, 0, size) == pytest.approx(base * size)
    assert opt.param_groups[0]["lr"] == pytest.approx(base * size)


def test_meters_and_accuracy():
    m = AverageMeter()
    m.update(1.0, 2)
    m.update(4.0, 1)
    assert m.avg == pytest.approx(2.0) and m.val == 4.0
    out = torch.tensor([[0.1, 0.9, 0.0, 0.0, 0.0, 0.0], [0.8, 0.1, 0.05, 0.03, 0.01, 0.01]])
    tgt = torch.tensor([1, 5])
    a1, a5 = accuracy(out, tgt, (1, 5))
    assert a1.item() == pytest.approx(50.0) and a5.item() == pytest.approx(50.0)
    assert top1_accuracy(out, tgt).item() == pytest.approx(0.5)
    met = Metric("x")
    met.update(torch.tensor(2.0))
    met.update(4.0)
    assert met.avg.item() == pytest.approx(3.0)


def test_dotenv_roundtrip(tmp_path):
    p = tmp_path / ".env"
    cfg.write_env_template(str(p), {"DATA": "/mnt/data"})
    vals = cfg.dotenv_values(str(p))
    assert vals["DATA"] == "/mnt/data" and "CLUSTER_MAX_NODES" in vals
    cfg.set_key(str(p), "CLUSTER_MAX_NODES", "4")
    cfg.set_key(str(p), "NEW_KEY", "x y")
    vals = cfg.dotenv_values(str(p))
    assert vals["CLUSTER_MAX_NODES"] == "4" and vals["NEW_KEY"] == "x y"
    sub = tmp_path / "a" / "b"
    sub.mkdir(parents=True)
    assert cfg.find_dotenv(str(sub)) == str(p)
    assert cfg.load_config(str(sub))["CLUSTER_MAX_NODES"] == "4"
    with pytest.raises(cfg.ConfigError):
        cfg.find_dotenv("/", filename=".definitely-missing-env", raise_error_if_not_found=True)


def test_dotenv_parsing_quotes_comments(tmp_path):
    p = tmp_path / ".env"
    p.write_text("# comment\nexport A=1\nB='two words'\nC=3 # trailing\n\nbad line\n")
    assert cfg.dotenv_values(str(p)) == {"A": "1", "B": "two words", "C": "3"}


def test_firelite_coercion():
    def main(training_data_path=None, use_gpu=False, epochs=5, base_lr=0.0125, model="resnet50"):
        return locals()

    out = firelite.Fire(main, ["--use_gpu", "True", "--epochs", "2", "--base-lr=0.1", "--model", "vgg16"])
    assert out == {"training_data_path": None, "use_gpu": True, "epochs": 2, "base_lr": 0.1, "model": "vgg16"}
    assert firelite.Fire(main, ["--use_gpu"])["use_gpu"] is True
    with pytest.raises(SystemExit):
        firelite.Fire(main, ["--nope", "1"])


def test_run_history(tmp_path):
    r = runs.Run("exp", root=str(tmp_path))
    r.tag("model", "resnet50")
    r.log_row("Training metrics", epoch=0, loss=1.5)
    r.rank_record(0, step=1, ms=2.0)
    r.complete()
    assert runs.list_experiments(str(tmp_path)) == ["exp"]
    listed = runs.list_runs("exp", str(tmp_path))
    assert listed[0]["status"] == "Completed" and listed[0]["tags"]["model"] == "resnet50"
    assert runs.read_metrics("exp", r.id, str(tmp_path))[0]["loss"] == 1.5
    os.environ["DDL_RUN_DIR"] = r.dir
    try:
        assert runs.Run.get_context().id == r.id
    finally:
        del os.environ["DDL_RUN_DIR"]


def test_metric_skips_collective_when_values_are_already_averaged():
    from distributeddeeplearning_b200.utils.meters import Metric

    m = Metric("loss")
    m.update(torch.tensor(2.0), averaged=True)
    m.update(torch.tensor(4.0), averaged=True)
    assert m._global and float(m.avg) == 3.0
    m.update(torch.tensor(6.0))                 # a local value: the mean must go through the allreduce again
    assert not m._global and float(m.avg) == 4.0


def test_fp8_parity_rule_tolerates_transient_noise_but_not_divergence():
    """tools/fp8_parity.py::judge — the acceptance rule of the fp8-vs-bf16 loss-curve test: the chaotic transient only has
    a blow-up guard, the settled last 30 % and the final window are tight, both arms must learn."""
    import importlib.util
    import math
    import os

    spec = importlib.util.spec_from_file_location(
        "fp8_parity", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fp8_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n = 200
    base = [7.0 + 1.5 * math.exp(-((i - 40) / 25.0) ** 2) - 0.2 * i / n for i in range(n)]      # overshoot, then decay
    wiggle = lambda amp, ph: [x * (1 + amp * math.sin(i / 7.0 + ph)) for i, x in enumerate(base)]
    b, b2 = wiggle(0.01, 0.0), wiggle(0.01, 1.0)
    # fp8 runs 9 % below bf16 during the overshoot only (what a real run looked like): accepted
    f = [x * (0.91 if 30 <= i < 90 else 1.0) for i, x in enumerate(wiggle(0.01, 2.0))]
    v = mod.judge(b, b2, f, 0.08)
    assert v["ok"] and 0.05 < v["gap_early"] < 0.12 and v["gap"] < 0.06 and v["learned"]
    # a curve that stays 12 % off in the settled part is rejected, and so is one that blows up early
    assert not mod.judge(b, b2, [x * (1.12 if i >= 130 else 1.0) for i, x in enumerate(f)], 0.08)["ok"]
    assert not mod.judge(b, b2, [x * (1.5 if 40 <= i < 80 else 1.0) for i, x in enumerate(f)], 0.08)["ok"]
    # not learning (flat loss) is rejected even if the curves agree
    flat = [7.0] * n
    assert not mod.judge(flat, flat, flat, 0.08)["ok"]
