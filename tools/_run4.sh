mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2x_tests.log; tail -4 gpurun_out/r2x_tests.log
MPH_IO_TRACE=1 timeout 900 python bench.py > gpurun_out/r2x_bench_exome.json 2> gpurun_out/r2x_bench_exome.err; echo bench rc=$?
tail -1 gpurun_out/r2x_bench_exome.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; e=j['e2e']
print('step %.3f ms' % j['ms_per_step'], 'e2e %.2f ms' % e['ms_per_step'], 'h2d', e['h2d_bytes_per_step'], 'frac %.3f' % r['frac'], r['kernel'], {k: round(v,3) for k,v in r['kernel_ms'].items()}, 'parity', j.get('parity_checked'))
print('e2e_files', j['e2e_files']['value'], j['e2e_files']['seconds'], j['e2e_files']['stages_ms'])"
grep "mph io" gpurun_out/r2x_bench_exome.err | tail -3
