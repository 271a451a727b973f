mkdir -p gpurun_out
for e in 0 1; do
if [ $e = 1 ]; then export MPH_REPLAY_STREAM=1; fi
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r24_bench_$e.log 2>&1; tail -1 gpurun_out/r24_bench_$e.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$e', j['ms_per_step'], j['roofline']['kernel_ms'])"
done
