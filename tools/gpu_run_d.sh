#!/bin/bash
# 1-GPU pass: A/B benches, fresh per-kernel launch list of one step, compute-sanitizer passes, full GPU test suite.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python tools/gpu_diag.py --groups conv_wgrad,conv_generic,sgd > gpurun_out/diag_d.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag_d.log
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>/dev/null | cut -c1-190 > gpurun_out/ab_default_$i.json
  DDL_DISABLE_STEM_TMA=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>/dev/null | cut -c1-190 > gpurun_out/ab_nostemtma_$i.json
done
DDL_WGRAD_SWAP=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>/dev/null | cut -c1-190 > gpurun_out/ab_noswap_1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/launches_run.log 2>&1
SAN_TIMEOUT=420 bash tools/sanitize.sh memcheck conv_generic > /dev/null 2>&1
SAN_TIMEOUT=300 bash tools/sanitize.sh racecheck bn > /dev/null 2>&1
SAN_TIMEOUT=300 bash tools/sanitize.sh synccheck conv_fwd > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
grep -E "FAIL|== group|rc=" gpurun_out/diag_d.log | head -20
for f in gpurun_out/ab_*.json; do echo "$f: $(cat $f | grep -o '"value": [0-9.]*, "unit": "images/sec", "n_gpus": 1, "steps": 30, "warmup": 5, "ms_per_step": [0-9.]*')"; done
wc -l gpurun_out/launches.csv
for f in gpurun_out/sanitize_*.log; do echo "--- $f"; grep -E "ERROR SUMMARY|rc=|Invalid|Race|hazard" $f | head -8; done
tail -4 gpurun_out/pytest_gpu.log
echo "total t=$(( $(date +%s) - T0 ))s"
