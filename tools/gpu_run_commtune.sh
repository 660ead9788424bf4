#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_run_commtune.sh N  — A/B of the bucket kernels' grid size / bucket size
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for cfg in "32 16" "8 16" "64 16" "32 4" "16 32"; do
  set -- $cfg; i=$((i+1))
  DDL_COMM_BLOCKS=$1 DDL_BUCKET_MB=$2 timeout 300 $TR --master-port $((29700+i)) bench.py --gpus $N --steps 30 --warmup 5 --no-e2e > gpurun_out/ct_${1}_${2}.json 2> gpurun_out/ct_${1}_${2}.err
  echo "blocks=$1 bucket_mb=$2: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ct_${1}_${2}.json) $(grep -o 'buckets=[0-9]*[^,]*' gpurun_out/ct_${1}_${2}.json | head -1)"
done
