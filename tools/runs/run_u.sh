mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 420 compute-sanitizer --target-processes all --tool racecheck --print-limit 200 $TR --master-port 29712 tools/comm_test.py --no-sweep > $O/san_race_N2.log 2>&1
echo "== racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|FAIL" $O/san_race_N2.log | tail -8; grep -c "ok\]" $O/san_race_N2.log
grep -E "Race reported|and Write|and Read" $O/san_race_N2.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head
