mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
for i in a b; do
for m in 1 2; do
DDL_PDL=$m timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e > $O/bench_pdl$m$i.json 2> $O/bench_pdl$m$i.err
cut -c1-230 $O/bench_pdl$m$i.json
done; done
