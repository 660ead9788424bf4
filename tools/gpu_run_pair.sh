#!/bin/bash
# cta_group::2 pair-MMA variant of the persistent kernel: numerics with the persistent path forced, then one A/B bench.
mkdir -p gpurun_out
T0=$(date +%s)
DDL_CONV_CLUSTER=2 DDL_CONV_PERSISTENT=2 timeout 90 python tools/gpu_diag.py --groups gemm > gpurun_out/diag_pair.log 2>&1
rc=$?; echo "gemm rc=$rc" >> gpurun_out/diag_pair.log
grep -E "FAIL|== group|rc=|rror" gpurun_out/diag_pair.log | head -12
if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=name,memory.used --format=csv,noheader; echo "pair mode failed; stopping"; exit 0; fi
DDL_CONV_CLUSTER=2 DDL_CONV_PERSISTENT=2 timeout 200 python tools/gpu_diag.py --groups conv_fwd,conv_dgrad,linear,conv_generic > gpurun_out/diag_pair2.log 2>&1
echo "rest rc=$?" >> gpurun_out/diag_pair2.log
grep -E "FAIL|== group|rc=" gpurun_out/diag_pair2.log | head -20
DDL_CONV_CLUSTER=2 timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e 2>gpurun_out/pair_on.err | cut -c1-200
DDL_CONV_CLUSTER=0 timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e 2>/dev/null | cut -c1-200
tail -2 gpurun_out/pair_on.err
echo "total t=$(( $(date +%s) - T0 ))s"
