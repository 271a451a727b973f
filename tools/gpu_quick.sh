mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "golden or c_abi or multi_device" 2>&1 | tail -3
python - <<'PY'
import time, os, sys
sys.path.insert(0, '.')
import microphaser_b200 as m
d = '/tmp/bamt'
m.synth_write_files(d, n_transcripts=2000, coverage=100.0)
for th in ("1", "8"):
    os.environ["MPH_IO_THREADS"] = th
    ctx = m.Context(0)
    t0 = time.perf_counter()
    ctx.run_somatic(d + '/reads.bam', d + '/ref.fa', d + '/variants.vcf', d + '/annotation.gtf', '/tmp/o.fa', '/tmp/o.tsv', '/tmp/o.n.fa')
    dt = time.perf_counter() - t0
    t = ctx.timing()
    print("io threads", th, "file-driven somatic, 2000 transcripts @100x: %.2f s wall; phase call %.1f ms; %d windows; %d records" % (dt, t["total_ms"], t["windows"], t["n_records"]))
    ctx.close()
PY
