#!/bin/bash
# Final 1-GPU validation: full GPU test suite, smoke, headline bench (both arms), fresh launch list + ncu captures.
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/final_ours_$i.json 2> gpurun_out/final_ours_$i.err; done
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/final_ref_1.json 2> gpurun_out/final_ref_1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --cuda-graph off > gpurun_out/launches_run.log 2>&1
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 8 -f -o gpurun_out/prof_conv \
      python tools/ncu_target.py conv > gpurun_out/prof_conv.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -c 4 -f -o gpurun_out/prof_wgrad \
      python tools/ncu_target.py conv > gpurun_out/prof_wgrad.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_act -s 3 -c 3 -f -o gpurun_out/prof_bn \
      python tools/ncu_target.py bn > gpurun_out/prof_bn.log 2>&1
fi
LB_SWEEP=0 timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_final.log 2>&1
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300; tail -1 gpurun_out/smoke.log
for f in gpurun_out/final_*.json; do echo "$f: $(python - <<PY
import json
try:
    d=json.loads([l for l in open("$f") if l.startswith("{")][-1])
    print(d.get("impl"), round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), (d.get("config") or {}).get("cuda_graph"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
wc -l gpurun_out/launches.csv; tail -2 gpurun_out/layer_bench_final.log | cut -c1-300
echo "total t=$(( $(date +%s) - T0 ))s"
