mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
timeout 900 python -m pytest tests/test_gpu.py -x -q -k "fp8_loss_curve or kernel_group and bn" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
cp gpurun_out/fp8_parity.json $O/ 2>/dev/null
timeout 600 python tools/fp8_parity.py --steps 200 --batch 32 > $O/fp8_parity.log 2>&1; tail -5 $O/fp8_parity.log
