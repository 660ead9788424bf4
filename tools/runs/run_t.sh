# final 2-GPU validation of the round-2 build: engine checks, model-level equivalence, bench with self-check, sanitizer
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 tools/comm_test.py --no-sweep --model-check > $O/comm_N2.log 2>&1
echo "== comm"; grep -E "ok\]|FAIL|EQUIV|ENGINE|equiv" $O/comm_N2.log | tail -12; tail -2 $O/comm_N2.log
timeout 600 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_N2.json 2> $O/bench_N2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_N1.json 2> $O/bench_N1.err
cut -c1-260 $O/bench_N2.json $O/bench_N1.json; tail -2 $O/bench_N2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t/bench_N2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('engine_check','step_equivalence','weight_checksum','e2e','gpu_launches')})
PY
timeout 420 compute-sanitizer --target-processes all --tool memcheck --print-limit 20 $TR --master-port 29711 tools/comm_test.py --no-sweep > $O/san_mem_N2.log 2>&1
echo "== memcheck rc=$?"; grep -E "ERROR SUMMARY|ok\]|FAIL" $O/san_mem_N2.log | tail -8
timeout 420 compute-sanitizer --target-processes all --tool racecheck --print-limit 20 $TR --master-port 29712 tools/comm_test.py --no-sweep > $O/san_race_N2.log 2>&1
echo "== racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|ok\]|FAIL" $O/san_race_N2.log | tail -8
