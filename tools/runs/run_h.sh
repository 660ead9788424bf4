mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 300 python tools/gpu_diag.py --groups fp8,elementwise,zoo --timeout 280 > $O/diag.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 2400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1
echo "== diag"; grep -E "FAIL|group |rc=" $O/diag.log | head
echo "== bench"; cut -c1-400 $O/bench.json
echo "== pytest"; tail -15 $O/pytest_gpu.log
