#!/bin/bash
# 1-GPU pass: stem TMA numerics (+ gather fallback), BN tune, bench, small-batch check, layer bench.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python tools/gpu_diag.py --groups conv_fwd,conv_wgrad,bn,model > gpurun_out/diag_c.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_c.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ours_c.json 2> gpurun_out/ours_c.err
echo "ours rc=$?" >> gpurun_out/ours_c.err
DDL_DISABLE_STEM_TMA=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/ours_c_nostemtma.json 2> gpurun_out/ours_c_nostemtma.err
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model resnet50 --batch-size 64 --num-iters 3 > gpurun_out/b64.log 2>&1
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model inception_v3 --batch-size 128 --num-iters 3 > gpurun_out/incep.log 2>&1
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model densenet121 --batch-size 128 --num-iters 3 > gpurun_out/dense.log 2>&1
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model vgg16 --batch-size 128 --num-iters 3 > gpurun_out/vgg.log 2>&1
LB_SWEEP=0 timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_c.log 2>&1
grep -E "FAIL|== group|rc=" gpurun_out/diag_c.log | head -40
cut -c1-330 gpurun_out/ours_c.json; tail -2 gpurun_out/ours_c.err
cut -c1-200 gpurun_out/ours_c_nostemtma.json
for f in b64 incep dense vgg; do echo "--- $f"; tail -4 gpurun_out/$f.log | cut -c1-200; done
tail -2 gpurun_out/layer_bench_c.log | cut -c1-300
echo "total t=$(( $(date +%s) - T0 ))s"
