mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r14_tests.log
tail -4 gpurun_out/r14_tests.log
for p in 1 0; do
MPH_REPLAY_PRIORITY=$p timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r14_bench_$p.log 2>&1; tail -1 gpurun_out/r14_bench_$p.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['ms_per_step'], j['e2e']['stages_ms'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r14_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r14_ncu.log 2>&1
grep -E "k_replay|k_window_hist\(" gpurun_out/r14_launches.csv | head -4 | awk -F'","' '{print $5, $NF}'
