#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_diag.py --groups elementwise,avgdebug,sgd,model > gpurun_out/diag2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/ours.json 2> gpurun_out/ours.err
echo "ours rc=$?" >> gpurun_out/ours.err
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
grep -E "FAIL|group|cos|worst" gpurun_out/diag2.log | tail -40; cat gpurun_out/ours.json; tail -3 gpurun_out/ours.err; tail -4 gpurun_out/layer_bench.log; tail -2 gpurun_out/smoke.log
