mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
for d in 2 3 4; do
  DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=$d timeout 900 python tools/gpu_diag.py --groups gemm,conv_fwd,conv_dgrad,conv_generic,linear --timeout 280 > gpurun_out/r2a/deep$d.log 2>&1
done
timeout 600 python tools/gpu_diag.py --groups gemm,conv_fwd,conv_dgrad,model --timeout 280 > gpurun_out/r2a/auto.log 2>&1
timeout 1200 python tools/layer_bench.py > gpurun_out/r2a/layer_bench.log 2>&1
cp gpurun_out/layer_bench.json gpurun_out/r2a/layer_bench.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench_new.json 2> gpurun_out/r2a/bench_new.err
DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/r2a/bench_old.json 2> gpurun_out/r2a/bench_old.err
DDL_CONV_AUTOTUNE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > gpurun_out/r2a/bench_policy.json 2> gpurun_out/r2a/bench_policy.err
tail -3 gpurun_out/r2a/deep2.log gpurun_out/r2a/deep3.log gpurun_out/r2a/deep4.log gpurun_out/r2a/auto.log
tail -4 gpurun_out/r2a/layer_bench.log | cut -c1-600
cat gpurun_out/r2a/bench_new.json gpurun_out/r2a/bench_old.json gpurun_out/r2a/bench_policy.json | cut -c1-400
