mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 300 python tools/gpu_diag.py --group fp8 > $O/fp8.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_bf16.csv python bench.py --steps 2 --warmup 4 --no-e2e --cuda-graph off > $O/launches_bf16_run.log 2>&1
DDL_PRECISION=fp8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_fp8.csv python bench.py --steps 2 --warmup 4 --no-e2e --cuda-graph off > $O/launches_fp8_run.log 2>&1
python tools/parse_launches.py $O/launches_bf16.csv > $O/step_launch_list_bf16.md
python tools/parse_launches.py $O/launches_fp8.csv > $O/step_launch_list_fp8.md
tail -3 $O/fp8.log
head -30 $O/step_launch_list_bf16.md; head -40 $O/step_launch_list_fp8.md
