mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r13_bench.log 2>&1; tail -1 gpurun_out/r13_bench.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['ms_per_step'], j['e2e']['stages_ms'])"
