#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tools/gpu_diag.py --groups conv_dgrad > gpurun_out/diag_i.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_i.log
DDL_FUSE_BN_REDUCE=1 timeout 600 python tools/gpu_diag.py --groups model,zoo > gpurun_out/diag_i2.log 2>&1
echo "diag2 rc=$?" >> gpurun_out/diag_i2.log
for i in 1 2; do
  DDL_FUSE_BN_REDUCE=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/i_fuse_$i.json 2> gpurun_out/i_fuse_$i.err
  DDL_FUSE_BN_REDUCE=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/i_nofuse_$i.json 2> gpurun_out/i_nofuse_$i.err
done
grep -E "FAIL|== group|rc=" gpurun_out/diag_i.log | head -20
grep -E "FAIL|== group|rc=|worst" gpurun_out/diag_i2.log | head -20
for f in gpurun_out/i_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), d["config"].get("cuda_graph"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
tail -3 gpurun_out/i_fuse_1.err
echo "total t=$(( $(date +%s) - T0 ))s"
