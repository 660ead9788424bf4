mkdir -p gpurun_out/r2c2
O=gpurun_out/r2c2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 tools/comm_test.py --no-sweep --model-check > $O/comm_N2.log 2>&1
timeout 600 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_N2.json 2> $O/bench_N2.err
export SAN_TIMEOUT=400
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file $O/sanitize_memcheck_comm_rank%p.log $TR --master-port 29709 tools/comm_test.py --no-sweep > $O/comm_memcheck.log 2>&1
echo "== comm"; grep -E "ok\]|FAIL|EQUIV|ENGINE" $O/comm_N2.log | head -30
echo "== bench"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c2/bench_N2.json') if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","engine_check","step_equivalence","weight_checksum")})
PY
echo "== memcheck"; tail -5 $O/comm_memcheck.log; for f in $O/sanitize_memcheck_comm_rank*.log; do tail -2 $f; done
