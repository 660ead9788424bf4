mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 900 python tools/fp8_parity.py --steps 200 --batch 32 --fp8-mode mxfp8 > $O/mxfp8_parity.log 2>&1
DDL_PRECISION=fp8 DDL_FP8_MX=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_mxfp8.json 2> $O/bench_mxfp8.err
tail -6 $O/mxfp8_parity.log; cut -c1-300 $O/bench_mxfp8.json; tail -3 $O/bench_mxfp8.err
