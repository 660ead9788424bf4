#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tools/gpu_diag.py --groups conv_dgrad,bn > gpurun_out/diag_j.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_j.log
DDL_FUSE_BN_REDUCE=1 timeout 600 python tools/gpu_diag.py --groups model,zoo > gpurun_out/diag_j2.log 2>&1
echo "diag2 rc=$?" >> gpurun_out/diag_j2.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/j_$name.json 2> gpurun_out/j_$name.err; }
run base_1 DDL_FUSE_BN_REDUCE=0 DDL_FUSE_STEM_POOL=0
run stem_1 DDL_FUSE_BN_REDUCE=0 DDL_FUSE_STEM_POOL=1
run both_1 DDL_FUSE_BN_REDUCE=1 DDL_FUSE_STEM_POOL=1
run bnr_1 DDL_FUSE_BN_REDUCE=1 DDL_FUSE_STEM_POOL=0
run both_2 DDL_FUSE_BN_REDUCE=1 DDL_FUSE_STEM_POOL=1
run stem_2 DDL_FUSE_BN_REDUCE=0 DDL_FUSE_STEM_POOL=1
grep -E "FAIL|== group|rc=" gpurun_out/diag_j.log | head -20
grep -E "FAIL|== group|rc=|worst" gpurun_out/diag_j2.log | head -20
for f in gpurun_out/j_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), d["config"].get("cuda_graph"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
tail -3 gpurun_out/j_both_1.err
echo "total t=$(( $(date +%s) - T0 ))s"
