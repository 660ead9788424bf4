mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 60 ./tools/umma_probe > $O/umma_probe.log 2>&1
timeout 300 python tools/variant_probe.py 256,56,128,1,1,0 256,14,256,3,1,1 1024,14,256,1,1,0 128,28,128,3,1,1 512,28,1024,1,2,0 > $O/probe.log 2>&1
timeout 600 python tools/gpu_diag.py --groups fp8,benchshape --timeout 280 > $O/fp8_benchshape.log 2>&1
timeout 900 python tools/gpu_diag.py --groups model,zoograd,graph --timeout 400 > $O/model.log 2>&1
timeout 1200 python tools/layer_bench.py > $O/layer_bench.log 2>&1
cp gpurun_out/layer_bench.json $O/layer_bench.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_new.json 2> $O/bench_new.err
DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_old.json 2> $O/bench_old.err
timeout 900 python tools/fp8_parity.py --steps 200 --batch 32 > $O/fp8_parity.log 2>&1
cp gpurun_out/fp8_parity.json $O/ 2>/dev/null
echo "== umma"; cat $O/umma_probe.log
echo "== probe"; cat $O/probe.log | tail -60
echo "== fp8/benchshape"; grep -E "FAIL|group|rc=" $O/fp8_benchshape.log | head -40
echo "== model"; grep -E "FAIL|group |rc=" $O/model.log | head -20
echo "== layer"; tail -3 $O/layer_bench.log | cut -c1-300
echo "== bench"; cut -c1-330 $O/bench_new.json $O/bench_old.json; tail -3 $O/bench_new.err
echo "== parity"; tail -8 $O/fp8_parity.log
