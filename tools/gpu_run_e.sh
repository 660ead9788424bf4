#!/bin/bash
# 1-GPU pass: zoo, A/B of the tuning hooks (full JSON incl. clocks), CUDA-graph replay at batch 256 and 64.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tools/gpu_diag.py --groups zoo > gpurun_out/diag_e.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_e.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
}
run default_1 X=1
run nobn256_1 DDL_CONV_BN256=0
run noswap_1 DDL_WGRAD_SWAP=0
run nostemtma_1 DDL_DISABLE_STEM_TMA=1
run default_2 X=1
run noswap_2 DDL_WGRAD_SWAP=0
run nostemtma_2 DDL_DISABLE_STEM_TMA=1
timeout 300 python bench.py --steps 30 --warmup 5 --cuda-graph on > gpurun_out/ab_graph_1.json 2> gpurun_out/ab_graph_1.err
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model resnet50 --batch-size 64 --num-iters 3 > gpurun_out/b64_eager.log 2>&1
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model resnet50 --batch-size 64 --num-iters 3 --cuda-graph > gpurun_out/b64_graph.log 2>&1
grep -E "FAIL|ok\]|== group|rc=" gpurun_out/diag_e.log | head -12
for f in gpurun_out/ab_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), d["config"].get("cuda_graph"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
tail -3 gpurun_out/ab_graph_1.err
for f in b64_eager b64_graph; do echo "--- $f"; tail -5 gpurun_out/$f.log | cut -c1-200; done
echo "total t=$(( $(date +%s) - T0 ))s"
