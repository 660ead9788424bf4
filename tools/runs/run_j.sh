mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 400 python tools/gpu_diag.py --groups fp8,bn --timeout 280 > $O/diag.log 2>&1
DDL_PRECISION=fp8 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_fp8.json 2> $O/bench_fp8.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_bf16.json 2> $O/bench_bf16.err
grep -E "FAIL|group |rc=" $O/diag.log | head; cut -c1-260 $O/bench_fp8.json $O/bench_bf16.json; tail -2 $O/bench_fp8.err
