# one-off measurement: pipeline timeline of the end-to-end call (stage hand-overs, residue per stage)
mkdir -p gpurun_out
MPH_TIMELINE=1 timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/tl_bench.log 2> gpurun_out/timeline.log
tail -1 gpurun_out/tl_bench.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('default', j['e2e']['ms_per_step'], j['e2e']['stages_ms'])"
