mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
G=gemm,conv_generic,conv_fwd,conv_dgrad,benchshape,fp8,bn
timeout 400 python tools/gpu_diag.py --groups $G > $O/diag_default.log 2>&1; tail -3 $O/diag_default.log
DDL_CONV_PERSISTENT=2 DDL_CONV_DEEP=0 DDL_CONV_AUTOTUNE=0 timeout 300 python tools/gpu_diag.py --groups gemm,conv_generic,conv_fwd,conv_dgrad,benchshape > $O/diag_persist.log 2>&1; tail -2 $O/diag_persist.log
for m in 2 3; do DDL_CONV_DEEP=$m DDL_CONV_AUTOTUNE=0 timeout 300 python tools/gpu_diag.py --groups gemm,conv_generic,conv_fwd,conv_dgrad,benchshape > $O/diag_deep$m.log 2>&1; tail -2 $O/diag_deep$m.log; done
DDL_CONV_CLUSTER=2 DDL_CONV_PERSISTENT=2 timeout 300 python tools/gpu_diag.py --groups conv_fwd,conv_dgrad > $O/diag_cl2.log 2>&1; tail -2 $O/diag_cl2.log
timeout 600 python tools/epi_probe.py > $O/epi_probe.log 2>&1
cat $O/epi_probe.log
