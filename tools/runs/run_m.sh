mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 900 python tools/layer_bench.py > $O/layer_bench.log 2>&1
cp gpurun_out/layer_bench.json $O/layer_bench.json
tail -4 $O/layer_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench1.json 2> $O/bench1.err
tail -1 $O/bench1.json
python - <<'PY'
import json
from distributeddeeplearning_b200.ops import native as nv
PY
