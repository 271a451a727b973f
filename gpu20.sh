mkdir -p gpurun_out
MPH_TIMELINE=1 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r20_bench.log 2> gpurun_out/r20_timeline.log
grep "\[mph\]" gpurun_out/r20_timeline.log | tail -12
tail -1 gpurun_out/r20_bench.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['e2e']['stages_ms'], j['e2e']['value'])"
