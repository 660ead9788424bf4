"""Install the reference for bench.py's `--impl reference` arm.

1. Try the prescribed offline pip install of /root/reference into baseline/_ref.  The reference is a
   cookiecutter template with no setup.py / pyproject.toml, so pip refuses ("not installable").
2. Fall back to copying the reference's workload scripts VERBATIM (byte-identical) into
   baseline/_ref/ so they can be executed unmodified.  baseline/_ref is git-ignored.
The outcome is printed and recorded in baseline/_ref/INSTALL.json (and summarised in DESIGN.md).
"""
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DDL_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
TEMPLATE = "{{cookiecutter.project_name}}"
SCRIPTS = ["PyTorch_benchmark/src/pytorch_synthetic_benchmark.py",
           "PyTorch_imagenet/src/imagenet_pytorch_horovod.py", "PyTorch_imagenet/src/timer.py",
           "PyTorch_imagenet/src/logging.conf", "PyTorch_hvd/src/imagenet_pytorch_horovod.py",
           "PyTorch_hvd/src/timer.py", "PyTorch_hvd/src/logging.conf"]


def main() -> int:
    os.makedirs(DST, exist_ok=True)
    record = {"reference_root": REF}
    if not os.path.isdir(REF):
        record["status"] = "reference tree not present"
        print(json.dumps(record))
        return 1
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation",
                          "--find-links", "/opt/wheelhouse", "--target", DST, REF], capture_output=True, text=True)
    record["pip_returncode"] = pip.returncode
    record["pip_tail"] = (pip.stdout + pip.stderr).strip().splitlines()[-1:] if (pip.stdout + pip.stderr).strip() else []
    copied = []
    for rel in SCRIPTS:
        src = os.path.join(REF, TEMPLATE, rel)
        if os.path.isfile(src):
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            copied.append(rel)
    record["copied_verbatim"] = copied
    record["status"] = "scripts copied verbatim; run against baseline/hvd_shim (horovod not installable)"
    with open(os.path.join(DST, "INSTALL.json"), "w") as f:
        json.dump(record, f, indent=1)
    print(json.dumps(record))
    return 0 if copied else 1


if __name__ == "__main__":
    sys.exit(main())
