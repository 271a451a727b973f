mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r25_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r25_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2
tail -3 gpurun_out/r25_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r25_bench.log 2>&1; tail -1 gpurun_out/r25_bench.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['ms_per_step'], j['e2e']['stages_ms'], j['e2e']['h2d_bytes_per_step'], j['e2e']['value'])"
