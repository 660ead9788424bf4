mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29703 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_N4.json 2> $O/bench_N4.err
cut -c1-300 $O/bench_N4.json; tail -2 $O/bench_N4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z/bench_N4.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','engine_check','engine_check_detail','step_equivalence','weight_checksum','e2e','clocks')})
PY
