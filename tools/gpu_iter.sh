# one iteration on the GPU box: the parity tests of the benched shapes + replay-heavy profiles, then the bench line under a few settings
mkdir -p gpurun_out
T=${TAG:-it}
timeout 900 python -m pytest tests/test_gpu_config_shapes.py tests/test_gpu_parity.py tests/test_gpu_peptides.py -m gpu -x -q -k "${TESTK:-config or shape or carry or geom or golden or fs}" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
summ() { tail -1 "$1" | python -c "
import json,sys
try:
    j=json.loads(sys.stdin.read()); r=j['roofline']; e=j['e2e']
    print('$2', 'step %.3f ms' % j['ms_per_step'], 'e2e %.2f ms' % e['ms_per_step'], 'frac %.3f' % r['frac'], {k: round(v,3) for k,v in r['kernel_ms'].items()}, 'parity', j.get('parity_checked'))
except Exception as ex: print('$2', 'FAILED', ex)
"; }
i=0
while IFS= read -r envs; do
  i=$((i+1))
  env $envs timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 5 ${BENCHFLAGS:---no-cpu-baseline} > gpurun_out/${T}_bench_$i.json 2> gpurun_out/${T}_bench_$i.err
  summ gpurun_out/${T}_bench_$i.json "[$envs]"
done <<< "${VARIANTS:-MPH_X=0}"
