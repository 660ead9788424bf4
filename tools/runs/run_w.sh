mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in a b; do
for m in 0 1; do
DDL_PDL=$m timeout 600 $TR --master-port 2970$m bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --no-selfcheck > $O/bench_N2_pdl$m$i.json 2> $O/bench_N2_pdl$m$i.err
echo "N2 pdl=$m: $(cut -c1-200 $O/bench_N2_pdl$m$i.json | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | tr '\n' ' ')"
done; done
for m in 0 1; do
DDL_PDL=$m CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_N1_gpu1_pdl$m.json 2> /dev/null
echo "N1 gpu1 pdl=$m: $(grep -o '"ms_per_step": [0-9.]*' $O/bench_N1_gpu1_pdl$m.json)"
DDL_PDL=$m CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_N1_gpu0_pdl$m.json 2> /dev/null
echo "N1 gpu0 pdl=$m: $(grep -o '"ms_per_step": [0-9.]*' $O/bench_N1_gpu0_pdl$m.json)"
done
