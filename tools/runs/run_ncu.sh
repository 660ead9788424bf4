mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|conv_wgrad" -c 24 -o $O/prof_convlong python tools/ncu_target.py convlong > $O/ncu_convlong.log 2>&1
ncu -i $O/prof_convlong.ncu-rep --page raw --csv > $O/prof_convlong_raw.csv 2>/dev/null
ls -la $O
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2n/prof_convlong_raw.csv')))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','dram__bytes_read.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes.sum','lts__t_sectors_srcunit_tex_op_read.sum','smsp__inst_executed.sum','sm__inst_executed_pipe_uniform.sum']
idx=[hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print(' | '.join((r[i][:70] if i==idx[0] else r[i]) for i in idx))
PY
