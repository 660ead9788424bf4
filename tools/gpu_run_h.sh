#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tools/gpu_diag.py --groups conv_dgrad,bn,model > gpurun_out/diag_h.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_h.log
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/h_graph_$i.json 2> gpurun_out/h_graph_$i.err; done
grep -E "FAIL|== group|rc=|worst" gpurun_out/diag_h.log | head -20
for f in gpurun_out/h_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), d["config"].get("cuda_graph"), round((d.get("e2e") or {}).get("value",0),1), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
tail -3 gpurun_out/h_graph_1.err
echo "total t=$(( $(date +%s) - T0 ))s"
