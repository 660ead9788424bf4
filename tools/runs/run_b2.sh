mkdir -p gpurun_out/r2b2
O=gpurun_out/r2b2
timeout 600 python tools/gpu_diag.py --groups conv_fwd,conv_dgrad,conv_wgrad,conv_generic,benchshape,model --timeout 280 > $O/diag.log 2>&1
LB_VARIANTS=1 timeout 1200 python tools/layer_bench.py > $O/layer_bench.log 2>&1
cp gpurun_out/layer_bench.json $O/layer_bench.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_new.json 2> $O/bench_new.err
timeout 900 python tools/fp8_parity.py --steps 200 --batch 32 > $O/fp8_parity.log 2>&1
DDL_PRECISION=fp8 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_fp8.json 2> $O/bench_fp8.err
echo "== diag"; grep -E "FAIL|group |rc=" $O/diag.log | head -30
echo "== layer"; tail -3 $O/layer_bench.log | cut -c1-300
echo "== bench"; cut -c1-330 $O/bench_new.json $O/bench_fp8.json; tail -3 $O/bench_new.err $O/bench_fp8.err
echo "== parity"; tail -16 $O/fp8_parity.log
