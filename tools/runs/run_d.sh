mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 300 python tools/equiv_probe.py resnet50 32 > $O/equiv_b32.log 2>&1
DDL_CONV_AUTOTUNE=0 timeout 300 python tools/equiv_probe.py resnet50 32 > $O/equiv_b32_notune.log 2>&1
DDL_ASYNC_WGRAD=0 DDL_CONV_AUTOTUNE=0 timeout 300 python tools/equiv_probe.py resnet50 32 > $O/equiv_b32_sync.log 2>&1
timeout 300 python tools/equiv_probe.py resnet18 32 > $O/equiv_r18.log 2>&1
LB_FP8=1 LB_VARIANTS=0 timeout 900 python tools/layer_bench.py > $O/layer_bench_fp8.log 2>&1
for f in $O/equiv_*.log; do echo "== $f"; head -20 $f; done
grep -o "^[0-9x>k -]*x[0-9]* fwd *[0-9.]*ms\|fp8_us.*" $O/layer_bench_fp8.log | paste - - | head -30
