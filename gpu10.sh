mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r10_tests.log
tail -8 gpurun_out/r10_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r10_bench.log 2>&1; tail -2 gpurun_out/r10_bench.log
