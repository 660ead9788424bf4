mkdir -p gpurun_out/r2ab
O=gpurun_out/r2ab
timeout 900 python -m pytest tests/test_gpu.py -x -q -k "smoke or bench_contract or imagenet_trainer or hvd_trainer or graph_replay or single_rank_local or native_module or (kernel_group and (sgd or graph or model))" > $O/pytest.log 2>&1
tail -6 $O/pytest.log
