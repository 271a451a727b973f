mkdir -p gpurun_out
for st in 8 10; do
MPH_STAGES=$st timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/e2e_$st.log 2> gpurun_out/e2e_$st.err; tail -1 gpurun_out/e2e_$st.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$st', j['e2e']['ms_per_step'], j['e2e']['stages_ms'])"
done
