#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_diag.py --groups gemm,conv_fwd,conv_dgrad,conv_wgrad,linear,model > gpurun_out/diag6.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag6.log
LB_SWEEP=0 timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench5.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ours5.json 2> gpurun_out/ours5.err
echo "ours rc=$?" >> gpurun_out/ours5.err
DDL_ASYNC_WGRAD=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/ours5_syncwgrad.json 2> gpurun_out/ours5_syncwgrad.err
grep -E "FAIL|group|worst|TIMEOUT|rc=" gpurun_out/diag6.log | tail -40; tail -3 gpurun_out/layer_bench5.log | cut -c1-300; cat gpurun_out/ours5.json | cut -c1-330; cat gpurun_out/ours5_syncwgrad.json | cut -c1-330; tail -3 gpurun_out/ours5.err
bash tools/gpu_profile.sh
