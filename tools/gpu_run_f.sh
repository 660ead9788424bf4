#!/bin/bash
# 1-GPU pass: full GPU test suite + headline bench x3 (graph replay default) + eager once.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
for i in 1 2 3; do timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/f_graph_$i.json 2> gpurun_out/f_graph_$i.err; done
timeout 300 python bench.py --steps 30 --warmup 5 --cuda-graph off > gpurun_out/f_eager_1.json 2> gpurun_out/f_eager_1.err
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
for f in gpurun_out/f_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), d["config"].get("cuda_graph"), round((d.get("e2e") or {}).get("value",0),1), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
echo "total t=$(( $(date +%s) - T0 ))s"
