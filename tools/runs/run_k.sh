mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 600 python tools/epi_probe.py > $O/epi_probe.log 2>&1
tail -40 $O/epi_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm" -c 4 -o $O/prof_convwide python tools/ncu_target.py convwide > $O/ncu_convwide.log 2>&1
ncu -i $O/prof_convwide.ncu-rep --page raw --csv > $O/prof_convwide_raw.csv 2>/dev/null
ncu -i $O/prof_convwide.ncu-rep --page details --csv > $O/prof_convwide_details.csv 2>/dev/null
ls -la $O
