mkdir -p gpurun_out
MPH_TIMELINE=1 MPH_STAGES=8 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r19_bench.log 2> gpurun_out/r19_timeline.log
grep "\[mph\]" gpurun_out/r19_timeline.log | tail -12
MPH_TIMELINE=1 MPH_STAGES=8 MPH_HOST_THREADS=14 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r19_bench14.log 2> gpurun_out/r19_timeline14.log
grep "\[mph\]" gpurun_out/r19_timeline14.log | tail -12
