mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 python tools/gpu_diag.py --groups fp8,bn,model --timeout 280 > $O/diag.log 2>&1
timeout 900 python tools/fp8_parity.py --steps 200 --batch 32 > $O/fp8_parity.log 2>&1
cp gpurun_out/fp8_parity.json $O/ 2>/dev/null
DDL_PRECISION=fp8 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_fp8.json 2> $O/bench_fp8.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_bf16.json 2> $O/bench_bf16.err
DDL_PRECISION=fp8 timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model resnet50 --batch-size 256 --num-iters 3 --num-batches-per-iter 10 --phase-times > $O/phase_fp8.log 2>&1
timeout 300 python -m distributeddeeplearning_b200.workloads.benchmark --model resnet50 --batch-size 256 --num-iters 3 --num-batches-per-iter 10 --phase-times > $O/phase_bf16.log 2>&1
bash tools/runs/run_san.sh > $O/san.log 2>&1
echo "== diag"; grep -E "FAIL|group |rc=|twin" $O/diag.log | head -30
echo "== parity"; tail -4 $O/fp8_parity.log
echo "== bench"; cut -c1-300 $O/bench_fp8.json $O/bench_bf16.json; tail -3 $O/bench_fp8.err
echo "== phases"; grep -h "Phase\|Total img" $O/phase_fp8.log $O/phase_bf16.log
echo "== sanitizer"; tail -40 $O/san.log
