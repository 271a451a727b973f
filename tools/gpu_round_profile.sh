# round-2 artefacts on one B200: tests, bench lines of every workload, the reference arm, ncu launch list + full captures, e2e timeline
mkdir -p gpurun_out
nproc > gpurun_out/r2_host_cores.txt
timeout 1800 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log
MPH_IO_TRACE=1 timeout 900 python bench.py > gpurun_out/r2_bench_exome.json 2> gpurun_out/r2_bench_exome.err; echo bench rc=$?
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo ref rc=$?
for w in chr22 hypermutated normal filter; do timeout 900 python bench.py --workload $w --steps 5 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo $w rc=$?; done
MPH_TIMELINE=1 timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu-baseline > /dev/null 2> gpurun_out/r2_e2e_timeline.txt
timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_bench_exome.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo launches rc=$?
# full capture of one steady-state step, every kernel on one stream (the side chain is serialised under ncu anyway)
MPH_SIDE_REPLAY=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_read_decode|k_side_decode|k_allele_call|k_replay|k_read_runs|k_window_hist|k_assemble|k_rc_merge|k_rc_emit|k_rc_count" -s 60 -c 12 -o gpurun_out/prof_r2 python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo full rc=$?
ls -la gpurun_out/ | tail -15
