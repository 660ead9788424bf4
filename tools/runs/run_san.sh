# compute-sanitizer over the kernels the round-1 passes skipped: wgrad, persistent, deep-ring (pair + single), fp8, comm
mkdir -p gpurun_out/r2s
export SAN_TIMEOUT=500
bash tools/sanitize.sh memcheck conv_wgrad > gpurun_out/r2s/wgrad_mem.txt 2>&1
bash tools/sanitize.sh racecheck conv_wgrad > gpurun_out/r2s/wgrad_race.txt 2>&1
DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=2 bash tools/sanitize.sh memcheck conv_dgrad > gpurun_out/r2s/deep_pair_mem.txt 2>&1
cp gpurun_out/sanitize_memcheck_conv_dgrad.log gpurun_out/r2s/sanitize_memcheck_deep_pair_conv_dgrad.log
DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=3 bash tools/sanitize.sh racecheck gemm > gpurun_out/r2s/deep_race.txt 2>&1
cp gpurun_out/sanitize_racecheck_gemm.log gpurun_out/r2s/sanitize_racecheck_deep_gemm.log
DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=0 DDL_CONV_PERSISTENT=2 bash tools/sanitize.sh racecheck conv_fwd > gpurun_out/r2s/persist_race.txt 2>&1
cp gpurun_out/sanitize_racecheck_conv_fwd.log gpurun_out/r2s/sanitize_racecheck_persistent_conv_fwd.log
bash tools/sanitize.sh memcheck fp8 > gpurun_out/r2s/fp8_mem.txt 2>&1
bash tools/sanitize.sh synccheck sgd > gpurun_out/r2s/sgd_sync.txt 2>&1
cp gpurun_out/sanitize_*conv_wgrad.log gpurun_out/sanitize_memcheck_fp8.log gpurun_out/sanitize_synccheck_sgd.log gpurun_out/r2s/ 2>/dev/null
for f in gpurun_out/r2s/*.txt; do echo "== $f"; tail -3 $f; done
