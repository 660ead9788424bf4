#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_diag.py --groups bn,model,conv_fwd > gpurun_out/diag4.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag4.log
timeout 1200 python tools/layer_bench.py > gpurun_out/layer_bench3.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ours3.json 2> gpurun_out/ours3.err
echo "ours rc=$?" >> gpurun_out/ours3.err
grep -E "FAIL|group|cos|worst|TIMEOUT|rc=" gpurun_out/diag4.log | tail -40; tail -4 gpurun_out/layer_bench3.log | cut -c1-400; cat gpurun_out/ours3.json; tail -3 gpurun_out/ours3.err
