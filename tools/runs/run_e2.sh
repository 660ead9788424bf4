mkdir -p gpurun_out/r2e2
O=gpurun_out/r2e2
timeout 600 python tools/gpu_diag.py --groups fp8,bn --timeout 280 > $O/diag.log 2>&1
DDL_PRECISION=fp8 timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e > $O/bench_fp8.json 2> $O/bench_fp8.err
timeout 900 python tools/fp8_parity.py --steps 200 --batch 32 > $O/fp8_parity.log 2>&1
FAKE_DATA_LENGTH=20480 timeout 500 python -m invoke pytorch-imagenet.submit.local.synthetic --epochs 1 --batch-size 256 --precision fp8 > $O/imagenet_fp8_N1.log 2>&1
echo "== diag"; grep -E "FAIL|group |rc=|twin" $O/diag.log | head -30
echo "== parity"; tail -3 $O/fp8_parity.log
echo "== bench"; cut -c1-300 $O/bench_fp8.json; tail -3 $O/bench_fp8.err
echo "== imagenet"; tail -12 $O/imagenet_fp8_N1.log
