# 8-GPU evidence run (charged 8x: keep it short)
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 200 $TR8 --master-port 29801 tools/comm_test.py --no-sweep --model-check > $O/comm_N8.log 2>&1
SWEEP_MAX=$((1<<26)) timeout 240 $TR4 --master-port 29802 tools/comm_test.py > $O/comm_N4.log 2>&1
cp gpurun_out/comm_sweep_N4.json $O/ 2>/dev/null
timeout 120 $TR8 --master-port 29803 tools/comm_timeline.py --steps 4 > $O/timeline_N8.log 2>&1
timeout 240 $TR8 --master-port 29804 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_N8.json 2> $O/bench_N8.err
timeout 200 $TR8 --master-port 29806 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-selfcheck --fp16-allreduce > $O/bench_N8_bf16wire.json 2> $O/bench_N8_bf16wire.err
for wire in "" "--fp16-allreduce"; do
timeout 200 $TR8 --master-port 29808 -m distributeddeeplearning_b200.workloads.benchmark --model vgg16 --batch-size 128 --num-iters 2 --num-warmup-batches 5 --cuda-graph $wire > $O/vgg16_N8_${wire:-fp32}.log 2>&1
done
timeout 200 $TR4 --master-port 29809 -m distributeddeeplearning_b200.workloads.benchmark --model vgg16 --batch-size 128 --num-iters 2 --num-warmup-batches 5 --cuda-graph > $O/vgg16_N4_fp32.log 2>&1
timeout 240 $TR8 --master-port 29810 bench.py --impl reference --model vgg16 --batch-size 128 --gpus 8 --steps 10 --warmup 3 > $O/ref_vgg16_N8.json 2> $O/ref_vgg16_N8.err
timeout 240 $TR8 --master-port 29811 -m distributeddeeplearning_b200.workloads.benchmark --model resnet152 --batch-size 128 --num-iters 2 --num-warmup-batches 5 --cuda-graph > $O/resnet152_N8.log 2>&1
FAKE_DATA_LENGTH=122880 timeout 300 python -m invoke pytorch-imagenet.submit.remote.synthetic --node-count 8 --epochs 1 --batch-size 256 --precision fp8 > $O/imagenet_fp8_N8.log 2>&1
echo "== comm8"; grep -E "ok\]|FAIL|EQUIV|ENGINE" $O/comm_N8.log | head -20
echo "== comm4"; grep -E "KiB fp32|ENGINE" $O/comm_N4.log | tail -14
echo "== timeline8"; tail -14 $O/timeline_N8.log
echo "== bench"; for f in bench_N8 bench_N8_bf16wire ref_vgg16_N8; do echo $f; python - "$O/$f.json" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","engine_check","unavailable")}, (d.get("step_equivalence") or {}), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("ERR", e)
PY
done
echo "== zoo"; grep -H "Total img/sec" $O/vgg16_*.log $O/resnet152_*.log
echo "== imagenet"; grep -H "Total images/sec\|Precision" $O/imagenet_*_N8.log
tail -3 $O/bench_N8.err $O/imagenet_fp8_N8.log
