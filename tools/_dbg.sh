mkdir -p gpurun_out /tmp/s
oracle/_build/mph_synth_files /tmp/s 1297088515 1000 100 1 1 0 0 > /dev/null
( oracle/_build/mph_oracle somatic /tmp/s/reads.bam -r /tmp/s/ref.fa -b /tmp/s/variants.vcf -t /tmp/s/o.tsv -n /tmp/s/o.n.fa < /tmp/s/annotation.gtf > /tmp/s/o.fa ) &
run() {
  env $1 microphaser_b200/_lib/microphaser somatic /tmp/s/reads.bam -r /tmp/s/ref.fa -b /tmp/s/variants.vcf -t /tmp/s/g.tsv -n /tmp/s/g.n.fa < /tmp/s/annotation.gtf > /tmp/s/g.fa 2> /tmp/s/g.err
  rc=$?
  wait
  if cmp -s /tmp/s/g.tsv /tmp/s/o.tsv && cmp -s /tmp/s/g.fa /tmp/s/o.fa; then same=SAME; else same=DIFF; fi
  echo "[$1] rc=$rc $same $(tail -1 /tmp/s/g.err | cut -c1-150)"
}
for v in MPH_X=0 MPH_X=0 MPH_SIDE_REPLAY=0 MPH_SIDE_REPLAY=k2 MPH_BUS_SPAN_BYTES=1 MPH_MERGE_CTAS=0 MPH_IO_THREADS=1 MPH_PACK_THREADS=1 MPH_K2B_MINB=10 "MPH_SIDE_REPLAY=0 MPH_PACK_THREADS=1" "MPH_SIDE_REPLAY=0 MPH_MERGE_CTAS=0"; do run "$v"; done 2>&1 | tee gpurun_out/r2y_dbg.log
