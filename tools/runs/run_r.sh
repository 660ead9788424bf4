mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
for m in 1 0 3 1; do
DDL_PDL=$m timeout 300 python tools/gpu_diag.py --groups bn > $O/bn_pdl$m.log 2>&1
echo "== PDL=$m"; grep -E "FAIL|checks passed" $O/bn_pdl$m.log | head -8
done
for h in 0 2000 0 2000; do
DDL_CONV_WAIT_HINT=$h timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_hint$h.json 2> $O/bench_hint$h.err
echo "hint $h"; cut -c1-230 $O/bench_hint$h.json
done
