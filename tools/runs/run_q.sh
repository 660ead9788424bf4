mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
