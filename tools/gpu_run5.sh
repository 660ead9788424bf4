#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_diag.py --groups conv_fwd,conv_dgrad,conv_wgrad,bn,model > gpurun_out/diag5.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag5.log
LB_SWEEP=0 timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench4.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ours4.json 2> gpurun_out/ours4.err
echo "ours rc=$?" >> gpurun_out/ours4.err
grep -E "FAIL|group|worst|TIMEOUT|rc=" gpurun_out/diag5.log | tail -40; tail -3 gpurun_out/layer_bench4.log | cut -c1-300; cat gpurun_out/ours4.json; tail -3 gpurun_out/ours4.err
