mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
timeout 120 python tools/gpu_diag.py --groups bn > $O/bn_$i.log 2>&1
echo "run $i: $(grep -E 'checks passed' $O/bn_$i.log)"; grep -E "FAIL" $O/bn_$i.log | head -6
done
