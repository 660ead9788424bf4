#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tools/gpu_diag.py --groups bn,model,zoo > gpurun_out/diag_g.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_g.log
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/g_graph_$i.json 2> gpurun_out/g_graph_$i.err; done
LB_SWEEP=0 timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_g.log 2>&1
grep -E "FAIL|== group|rc=|worst" gpurun_out/diag_g.log | head -20
for f in gpurun_out/g_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), d["config"].get("cuda_graph"), round((d.get("e2e") or {}).get("value",0),1), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
tail -3 gpurun_out/g_graph_1.err
tail -2 gpurun_out/layer_bench_g.log | cut -c1-300
echo "total t=$(( $(date +%s) - T0 ))s"
