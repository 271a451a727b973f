mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 4 --no-cpu-baseline > gpurun_out/r26_$1.log 2>&1; tail -1 gpurun_out/r26_$1.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$1', j['e2e']['ms_per_step'], j['e2e']['stages_ms'])"; }
run base
MALLOC_MMAP_THRESHOLD_=33554432 MALLOC_TRIM_THRESHOLD_=2147483648 MALLOC_TOP_PAD_=268435456 run tuned
MALLOC_ARENA_MAX=64 MALLOC_MMAP_THRESHOLD_=33554432 MALLOC_TRIM_THRESHOLD_=2147483648 run tuned_arena
MPH_HOST_THREADS=8 run t8
MPH_HOST_THREADS=32 run t32
