#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/gpu_diag.py > gpurun_out/diag3.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag3.log
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ours2.json 2> gpurun_out/ours2.err
echo "ours rc=$?" >> gpurun_out/ours2.err
grep -E "FAIL|group|cos|worst|TIMEOUT|rc=" gpurun_out/diag3.log | tail -50; tail -5 gpurun_out/layer_bench2.log; cat gpurun_out/ours2.json; tail -3 gpurun_out/ours2.err
