python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --no-cpu-baseline > gpurun_out/bench_exome_b.json 2> gpurun_out/bench_exome_b.err; echo rc=$?; python -c "
import json; d=json.loads(open('gpurun_out/bench_exome_b.json').read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['e2e']['stages_ms'])"
tail -3 gpurun_out/bench_exome_b.err
