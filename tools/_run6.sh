TAG=r2aa TESTK="config or shape or pipelined or carry" VARIANTS=$'MPH_X=0\nMPH_SIDE_REPLAY=0\nMPH_SIDE_REPLAY=k2' bash tools/gpu_iter.sh
MPH_IO_TRACE=1 timeout 900 python bench.py --steps 5 > gpurun_out/r2aa_bench_full.json 2> gpurun_out/r2aa_bench_full.err; echo bench rc=$?
tail -1 gpurun_out/r2aa_bench_full.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']; e=j['e2e']
print('step %.3f ms' % j['ms_per_step'], 'e2e %.2f ms' % e['ms_per_step'], 'frac %.3f' % r['frac'], r['kernel'], {k: round(v,3) for k,v in r['kernel_ms'].items()}, 'parity', j.get('parity_checked'))
print('e2e_files', j['e2e_files']['value'], j['e2e_files']['seconds'], j['e2e_files']['stages_ms'])"
grep "mph io" gpurun_out/r2aa_bench_full.err | tail -3
mkdir -p /tmp/s; oracle/_build/mph_synth_files /tmp/s 1297088515 1000 100 1 1 0 0 > /dev/null
for v in MPH_PARSE_THREADS=2 MPH_PARSE_THREADS=4 MPH_PARSE_THREADS=8 "MPH_PARSE_THREADS=4 MPH_IO_THREADS=12"; do
  for rep in 1 2; do env $v MPH_IO_TRACE=1 microphaser_b200/_lib/microphaser somatic /tmp/s/reads.bam -r /tmp/s/ref.fa -b /tmp/s/variants.vcf -t /tmp/s/g.tsv -n /tmp/s/g.n.fa < /tmp/s/annotation.gtf 2>&1 > /tmp/s/g.fa | grep "mph io" | sed "s/^/[$v] /"; done
done
nproc
