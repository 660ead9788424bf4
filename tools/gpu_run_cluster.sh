#!/bin/bash
# A/B of the CTA-pair (TMA multicast) persistent kernel: numerics with the persistent path forced, then bench.
mkdir -p gpurun_out
T0=$(date +%s)
DDL_CONV_CLUSTER=1 DDL_CONV_PERSISTENT=2 timeout 300 python tools/gpu_diag.py --groups gemm,conv_fwd,conv_dgrad,linear,conv_generic > gpurun_out/diag_cl.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_cl.log
DDL_CONV_CLUSTER=1 timeout 300 python tools/gpu_diag.py --groups model > gpurun_out/diag_cl2.log 2>&1
echo "diag2 rc=$?" >> gpurun_out/diag_cl2.log
for i in 1 2; do
  DDL_CONV_CLUSTER=1 timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/cl_on_$i.json 2> gpurun_out/cl_on_$i.err
  DDL_CONV_CLUSTER=0 timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/cl_off_$i.json 2> gpurun_out/cl_off_$i.err
done
grep -E "FAIL|== group|rc=" gpurun_out/diag_cl.log | head -30
grep -E "FAIL|== group|rc=|worst" gpurun_out/diag_cl2.log | head
for f in gpurun_out/cl_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), d["config"].get("cuda_graph"))
except Exception as e:
    print("ERR", e)
PY
)"; done
tail -3 gpurun_out/cl_on_1.err
echo "total t=$(( $(date +%s) - T0 ))s"
