mkdir -p gpurun_out /tmp/s
oracle/_build/mph_synth_files /tmp/s 1297088515 1000 100 1 1 0 0 > /dev/null
oracle/_build/mph_oracle somatic /tmp/s/reads.bam -r /tmp/s/ref.fa -b /tmp/s/variants.vcf -t /tmp/s/o.tsv -n /tmp/s/o.n.fa < /tmp/s/annotation.gtf > /tmp/s/o.fa &
timeout 900 python -m pytest tests/test_gpu_config_shapes.py tests/test_gpu_parity.py -m gpu -x -q -k "config or golden" > gpurun_out/r2ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ab_tests.log; tail -3 gpurun_out/r2ab_tests.log
wait
for v in MPH_X=0 MPH_X=1 MPH_GPU_INFLATE=0 MPH_GPU_INFLATE_BLOCKS=1024 MPH_GPU_INFLATE_BLOCKS=8192; do
  env $v MPH_IO_TRACE=1 MPH_TIMELINE=1 microphaser_b200/_lib/microphaser somatic /tmp/s/reads.bam -r /tmp/s/ref.fa -b /tmp/s/variants.vcf -t /tmp/s/g.tsv -n /tmp/s/g.n.fa < /tmp/s/annotation.gtf > /tmp/s/g.fa 2> /tmp/s/g.err; rc=$?
  if cmp -s /tmp/s/g.tsv /tmp/s/o.tsv && cmp -s /tmp/s/g.fa /tmp/s/o.fa && cmp -s /tmp/s/g.n.fa /tmp/s/o.n.fa; then same=SAME; else same=DIFF; fi
  echo "[$v] rc=$rc $same"; grep "mph io" /tmp/s/g.err | sed 's/^/    /'
done
MPH_IO_TRACE=1 timeout 900 python bench.py --steps 5 > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo bench rc=$?
tail -1 gpurun_out/r2ab_bench.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; e=j['e2e']
print('step %.3f ms' % j['ms_per_step'], 'e2e %.2f ms' % e['ms_per_step'], 'frac %.3f' % r['frac'], 'launches', j['gpu_launches'], 'parity', j.get('parity_checked'))
print('e2e_files', j['e2e_files']['value'], j['e2e_files']['seconds'], j['e2e_files']['stages_ms'])"
grep "mph io" gpurun_out/r2ab_bench.err | tail -6
timeout 300 python bench.py --workload hypermutated --steps 5 --no-cpu-baseline | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; print('hyper step %.3f' % j['ms_per_step'], 'serial %.3f' % r['serial_chain_ms'], {k: round(v,3) for k,v in r['kernel_ms_in_timed_loop'].items()})"
