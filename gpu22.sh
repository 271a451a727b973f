mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r22_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r22_tests.log
tail -3 gpurun_out/r22_tests.log
MPH_SYNTH_MODE=1 timeout 900 python bench.py --workload chr22 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/r22_bench_normal.log 2>&1; tail -1 gpurun_out/r22_bench_normal.log | cut -c1-1800
