mkdir -p gpurun_out/r2aa
O=gpurun_out/r2aa
export DDL_NATIVE_HOOKS=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29701 tools/comm_test.py --no-sweep --model-check > $O/comm_N2.log 2>&1
echo "== comm rc=$?"; grep -E "FAIL|EQUIV|ENGINE|Error|error" $O/comm_N2.log | tail -8; grep -c "ok\]" $O/comm_N2.log
timeout 500 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_N2.json 2> $O/bench_N2.err
echo "== bench rc=$?"; cut -c1-200 $O/bench_N2.json; tail -3 $O/bench_N2.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2aa/bench_N2.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('engine_check','step_equivalence','weight_checksum','gpu_launches','final_loss')})
except Exception as e: print('no json', e)
PY
