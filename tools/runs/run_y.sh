mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_N1.json 2> $O/bench_N1.err
cat $O/bench_N1.json | cut -c1-1200; tail -2 $O/bench_N1.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python -c "
import sys; sys.path.insert(0,'.')
import torch, bench
s=bench.ClockSampler(0); print('sampler target', s.gpu)"
timeout 300 python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
