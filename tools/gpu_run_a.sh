#!/bin/bash
# 1-GPU validation pass: GPU test suite, smoke, headline bench (device + e2e), model zoo, ncu captures of the conv kernels.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ours_a.json 2> gpurun_out/ours_a.err
echo "ours rc=$?" >> gpurun_out/ours_a.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 8 -f -o gpurun_out/prof_conv \
    python tools/ncu_target.py conv > gpurun_out/prof_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -c 4 -f -o gpurun_out/prof_wgrad \
    python tools/ncu_target.py conv > gpurun_out/prof_wgrad.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-700 gpurun_out/ours_a.json; tail -2 gpurun_out/ours_a.err
ls -la gpurun_out/*.ncu-rep
echo "total t=$(( $(date +%s) - T0 ))s"
