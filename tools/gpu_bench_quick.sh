mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bq.log 2>&1; tail -1 gpurun_out/bq.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['kernel_ms'])"
