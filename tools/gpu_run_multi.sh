#!/bin/bash
# usage: bash tools/gpu_run_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_N$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SWEEP_MAX=${SWEEP_MAX:-268435456} timeout 900 $TR --master-port 29611 tools/comm_test.py > gpurun_out/comm_N$N.log 2>&1
echo "comm rc=$?" >> gpurun_out/comm_N$N.log
timeout 600 $TR --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/ours_N$N.json 2> gpurun_out/ours_N$N.err
echo "ours rc=$?" >> gpurun_out/ours_N$N.err
timeout 600 $TR --master-port 29613 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/ref_N$N.json 2> gpurun_out/ref_N$N.err
timeout 600 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 5 --fp16-allreduce --no-e2e > gpurun_out/ours_bf16wire_N$N.json 2> gpurun_out/ours_bf16wire_N$N.err
grep -E "ok\]|FAIL|ENGINE|rc=|Error|error" gpurun_out/comm_N$N.log | tail -30
tail -12 gpurun_out/comm_N$N.log
cat gpurun_out/ours_N$N.json gpurun_out/ref_N$N.json gpurun_out/ours_bf16wire_N$N.json; tail -5 gpurun_out/ours_N$N.err
