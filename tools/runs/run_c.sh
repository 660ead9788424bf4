# 2-GPU validation of the rewritten fused allreduce+SGD kernel
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 tools/comm_test.py --no-sweep --model-check > $O/comm_N2.log 2>&1
timeout 300 $TR --master-port 29702 tools/comm_timeline.py > $O/timeline_N2.log 2>&1
timeout 600 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_N2.json 2> $O/bench_N2.err
timeout 600 $TR --master-port 29704 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-selfcheck --fp16-allreduce > $O/bench_N2_bf16wire.json 2> $O/bench_N2_bf16wire.err
DDL_BENCH_SAME_DATA=1 timeout 600 $TR --master-port 29705 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-selfcheck > $O/bench_N2_samedata.json 2> $O/bench_N2_samedata.err
for blocks in 16 64; do
DDL_COMM_BLOCKS=$blocks timeout 600 $TR --master-port 29706 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-selfcheck > $O/bench_N2_b$blocks.json 2> $O/bench_N2_b$blocks.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_N1.json 2> $O/bench_N1.err
for wire in "" "--fp16-allreduce"; do
timeout 600 $TR --master-port 29707 -m distributeddeeplearning_b200.workloads.benchmark --model vgg16 --batch-size 128 --num-iters 3 --cuda-graph $wire > $O/vgg16_N2_${wire:-fp32}.log 2>&1
done
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model vgg16 --batch-size 128 --num-iters 3 --cuda-graph > $O/vgg16_N1.log 2>&1
echo "== comm"; grep -E "ok\]|FAIL|EQUIV|ENGINE" $O/comm_N2.log | head -30; tail -3 $O/comm_N2.log
echo "== timeline"; tail -16 $O/timeline_N2.log
echo "== bench"; for f in $O/bench_N1.json $O/bench_N2.json $O/bench_N2_bf16wire.json $O/bench_N2_samedata.json $O/bench_N2_b16.json $O/bench_N2_b64.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print({k:d.get(k) for k in ("value","ms_per_step","final_loss","engine_check","step_equivalence","weight_checksum")}, (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("ERR", e)
PY
done
echo "== vgg"; grep -h "Total img/sec\|Img/sec per" $O/vgg16_*.log
tail -3 $O/bench_N2.err
