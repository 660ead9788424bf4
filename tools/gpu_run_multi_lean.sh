#!/bin/bash
# usage: bash tools/gpu_run_multi_lean.sh N [ref]  (under gpurun --gpus N): engine checks (no sweep), headline bench with
# graph replay and eager, hvd trainer smoke (piggy-backed metrics), optional reference arm.
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$2" = "benchonly" ]; then     # just the headline line (one scaling point)
  timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/oursg_N$N.json 2> gpurun_out/oursg_N$N.err
  echo "ours(graph) rc=$?" >> gpurun_out/oursg_N$N.err
  grep -o '"value": [0-9.]*, "unit": "images/sec", "n_gpus": [0-9]*, "steps": 30, "warmup": 5, "ms_per_step": [0-9.]*' gpurun_out/oursg_N$N.json
  tail -2 gpurun_out/oursg_N$N.err
  exit 0
fi
timeout 600 $TR --master-port 29611 tools/comm_test.py --no-sweep > gpurun_out/commlean_N$N.log 2>&1
echo "comm rc=$?" >> gpurun_out/commlean_N$N.log
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/oursg_N$N.json 2> gpurun_out/oursg_N$N.err
echo "ours(graph) rc=$?" >> gpurun_out/oursg_N$N.err
timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 30 --warmup 5 --cuda-graph off --no-e2e > gpurun_out/ourse_N$N.json 2> gpurun_out/ourse_N$N.err
echo "ours(eager) rc=$?" >> gpurun_out/ourse_N$N.err
timeout 600 $TR --master-port 29615 -m distributeddeeplearning_b200.workloads.hvd_imagenet --epochs 1 --batch-size 32 --model resnet50 \
    --synthetic-length 256 --checkpoint-format /tmp/ck-{epoch}.pth.tar --log-dir /tmp/hvdlogs > gpurun_out/hvd_N$N.log 2>&1
echo "hvd rc=$?" >> gpurun_out/hvd_N$N.log
if [ "$2" = "ref" ]; then
  timeout 600 $TR --master-port 29614 bench.py --impl reference --gpus $N --steps 30 --warmup 5 > gpurun_out/ref2_N$N.json 2> gpurun_out/ref2_N$N.err
fi
grep -E "ok\]|FAIL|ENGINE|rc=|Error|error" gpurun_out/commlean_N$N.log | tail -12
for f in gpurun_out/oursg_N$N.json gpurun_out/ourse_N$N.json gpurun_out/ref2_N$N.json; do [ -f $f ] && echo "$f: $(python - <<PY
import json
try:
    d=json.loads([l for l in open("$f") if l.startswith("{")][-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), d.get("clocks"), (d.get("config") or {}).get("cuda_graph"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", e)
PY
)"; done
tail -3 gpurun_out/oursg_N$N.err; tail -4 gpurun_out/hvd_N$N.log | cut -c1-200
echo "total t=$(( $(date +%s) - T0 ))s"
