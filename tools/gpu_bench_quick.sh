mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bq.log 2>&1; tail -1 gpurun_out/bq.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['roofline']['survey_model'], j['roofline']['frac'])"
