mkdir -p gpurun_out
for i in 1 2; do
MPH_TIMELINE=1 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 4 --no-cpu-baseline > gpurun_out/e2e_$i.log 2> gpurun_out/e2e_$i.err; tail -1 gpurun_out/e2e_$i.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['e2e']['ms_per_step'], j['e2e']['stages_ms'])"
done
grep "\[mph\]" gpurun_out/e2e_2.err | tail -12
