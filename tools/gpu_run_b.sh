#!/bin/bash
# 1-GPU pass: generic-shape conv/BN numerics, regression groups, zoo, bench, layer bench.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python tools/gpu_diag.py --groups conv_generic > gpurun_out/diag_generic.log 2>&1
echo "generic rc=$?" >> gpurun_out/diag_generic.log
timeout 900 python tools/gpu_diag.py --groups gemm,conv_fwd,conv_dgrad,conv_wgrad,linear,bn,elementwise,sgd > gpurun_out/diag_b.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_b.log
timeout 900 python tools/gpu_diag.py --groups zoo,zoograd,model > gpurun_out/diag_zoo.log 2>&1
echo "zoo rc=$?" >> gpurun_out/diag_zoo.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ours_b.json 2> gpurun_out/ours_b.err
echo "ours rc=$?" >> gpurun_out/ours_b.err
LB_SWEEP=0 timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_b.log 2>&1
grep -E "FAIL|== group|rc=" gpurun_out/diag_generic.log | head -60
grep -E "FAIL|== group|rc=" gpurun_out/diag_b.log | head -30
grep -E "FAIL|ok\]|== group|rc=|worst|cos" gpurun_out/diag_zoo.log | head -30
cut -c1-400 gpurun_out/ours_b.json; tail -2 gpurun_out/ours_b.err
tail -2 gpurun_out/layer_bench_b.log | cut -c1-300
echo "total t=$(( $(date +%s) - T0 ))s"
