mkdir -p gpurun_out/r2n2
O=gpurun_out/r2n2
timeout 500 python tools/gpu_diag.py --groups graph,conv_fwd,conv_dgrad,bn,model,gemm > $O/diag.log 2>&1; grep -E "FAILED|rc=" $O/diag.log | tail -8
timeout 300 python -m pytest tests/test_gpu.py -x -q -k "graph_replay or smoke or whole_model" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
DDL_PDL=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_pdl0.json 2> $O/bench_pdl0.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_pdl1.json 2> $O/bench_pdl1.err
DDL_PDL=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_pdl0b.json 2> $O/bench_pdl0b.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_pdl1b.json 2> $O/bench_pdl1b.err
cut -c1-230 $O/bench_pdl0.json $O/bench_pdl1.json $O/bench_pdl0b.json $O/bench_pdl1b.json; tail -2 $O/bench_pdl1.err
