mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r18_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r18_tests.log
tail -5 gpurun_out/r18_tests.log
for st in 8 1; do
MPH_STAGES=$st timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r18_bench_$st.log 2>&1; tail -1 gpurun_out/r18_bench_$st.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['kernel_ms'], j['e2e']['ms_per_step'], j['e2e']['stages_ms'], j['e2e']['value'])"
done
