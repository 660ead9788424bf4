#!/bin/bash
# first GPU bring-up: kernel diagnostics, reference-arm baseline, our arm
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count)" > gpurun_out/dev.txt 2>&1
timeout 1500 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ref_b256.json 2> gpurun_out/ref_b256.err
timeout 300 python bench.py --impl reference --batch-size 64 --steps 30 --warmup 5 > gpurun_out/ref_b64.json 2> gpurun_out/ref_b64.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/ours.json 2> gpurun_out/ours.err
echo "ours rc=$?" >> gpurun_out/ours.err
tail -5 gpurun_out/diag.log; cat gpurun_out/ref_b256.json gpurun_out/ours.json; tail -3 gpurun_out/ours.err
