mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "golden or c_abi or multi_device" 2>&1 | tail -2
python - <<'PY'
import time, os, sys
sys.path.insert(0, '.')
import microphaser_b200 as m
d = '/tmp/bamt'
m.synth_write_files(d, n_transcripts=2000, coverage=100.0)
for th, pk in (("1", "1"), ("8", "4"), ("8", "4")):
    os.environ["MPH_IO_THREADS"] = th; os.environ["MPH_PACK_THREADS"] = pk
    ctx = m.Context(0)
    t0 = time.perf_counter()
    ctx.run_somatic(d + '/reads.bam', d + '/ref.fa', d + '/variants.vcf', d + '/annotation.gtf', '/tmp/o%s.fa' % th, '/tmp/o%s.tsv' % th, '/tmp/o%s.n.fa' % th)
    dt = time.perf_counter() - t0
    t = ctx.timing()
    print("io threads", th, "pack shards", pk, "file-driven somatic, 2000 transcripts @100x: %.2f s wall; phase calls %.1f ms; %d windows; %d records" % (dt, t["total_ms"], t["windows"], t["n_records"]))
    ctx.close()
print("identical outputs:", all(open('/tmp/o1' + e, 'rb').read() == open('/tmp/o8' + e, 'rb').read() for e in ('.fa', '.tsv', '.n.fa')))
PY
