"""GPU tests (B200): every native kernel vs a plain PyTorch fp32 reference of the same op, the fused engine,
whole-model gradients, and the bench/smoke entry points.  Each group runs in a subprocess (tools/gpu_diag.py) so a
device-side trap in one kernel cannot poison the CUDA context of the others."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _group(name, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_diag.py"), "--group", name], cwd=ROOT,
                       capture_output=True, text=True, timeout=timeout)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert "FAIL" not in r.stdout, tail


@pytest.mark.parametrize("group", ["elementwise", "gemm", "conv_fwd", "conv_dgrad", "conv_wgrad", "linear", "bn", "sgd", "zoo", "zoograd", "conv_generic"])
def test_kernel_group(group):
    _group(group)


def test_whole_model_gradients():
    _group("model", timeout=900)


def test_native_module_loaded_and_counts_launches():
    code = ("import torch, sys; sys.path.insert(0, %r)\n"
            "from distributeddeeplearning_b200 import _ext\n"
            "from distributeddeeplearning_b200.ops import native as nv\n"
            "n0 = _ext.launch_count(); x = nv.philox_images(2, 8, 8, 1); torch.cuda.synchronize()\n"
            "assert _ext.launch_count() == n0 + 1\n"
            "import distributeddeeplearning_b200._C as m; print(m.__file__)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().endswith("_C.so")


def test_smoke_entry():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "__graft_entry__.py"), "smoke"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "smoke ok" in r.stdout


def test_bench_contract_small():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--batch-size", "32"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches"):
        assert key in out, key
    assert out["value"] > 0 and out["gpu_launches"] > 0 and out["e2e"]["value"] > 0
    assert out["e2e"]["h2d_bytes_per_step"] == 32 * 224 * 224 * 3 + 32 * 8
