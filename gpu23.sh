mkdir -p gpurun_out
for st in 8 12 16 24; do
MPH_STAGES=$st timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r23_bench_$st.log 2>&1; tail -1 gpurun_out/r23_bench_$st.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$st', j['e2e']['ms_per_step'], j['e2e']['stages_ms'])"
done
