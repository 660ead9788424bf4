# after the epilogue rewrite: racecheck / memcheck of the statistics-carrying conv kernels, then the whole GPU suite
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
export SAN_TIMEOUT=400
DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=0 DDL_CONV_PERSISTENT=2 bash tools/sanitize.sh racecheck conv_fwd > $O/persist_race.txt 2>&1
cp gpurun_out/sanitize_racecheck_conv_fwd.log $O/sanitize_racecheck_persistent_conv_fwd.log
DDL_CONV_AUTOTUNE=0 DDL_CONV_DEEP=3 bash tools/sanitize.sh racecheck conv_fwd > $O/deep_race.txt 2>&1
cp gpurun_out/sanitize_racecheck_conv_fwd.log $O/sanitize_racecheck_deep_conv_fwd.log
bash tools/sanitize.sh memcheck conv_fwd > $O/fwd_mem.txt 2>&1
cp gpurun_out/sanitize_memcheck_conv_fwd.log $O/sanitize_memcheck_conv_fwd.log
for f in $O/*.txt; do echo "== $f"; tail -3 $f; done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
tail -5 $O/pytest_gpu.log
