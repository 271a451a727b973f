mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests.log; tail -4 gpurun_out/r2_tests.log
MPH_IO_TRACE=1 timeout 600 python bench.py > gpurun_out/r2_bench_exome.json 2> gpurun_out/r2_bench_exome.err; echo bench rc=$?
for w in normal hypermutated chr22; do timeout 300 python bench.py --workload $w --steps 5 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo $w rc=$?; done
MPH_TIMELINE=1 timeout 300 python bench.py --steps 2 --warmup 3 --e2e-steps 2 --no-cpu-baseline > /dev/null 2> gpurun_out/r2_e2e_timeline.txt
