#!/bin/bash
# Profiling pass (1 GPU): launch list of one training step + full ncu captures of the hot kernels.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 2 -c 4 -f -o gpurun_out/prof_conv \
    python tools/ncu_target.py conv > gpurun_out/prof_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kernel -s 1 -c 2 -f -o gpurun_out/prof_wgrad \
    python tools/ncu_target.py conv > gpurun_out/prof_wgrad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_act -s 3 -c 3 -f -o gpurun_out/prof_bn \
    python tools/ncu_target.py bn > gpurun_out/prof_bn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sgd -s 3 -c 1 -f -o gpurun_out/prof_sgd \
    python tools/ncu_target.py sgd > gpurun_out/prof_sgd.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/launches_run.log; wc -l gpurun_out/launches.csv
