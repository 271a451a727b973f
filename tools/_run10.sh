TAG=r2ac TESTK="config or shape or pipelined or multi_device" VARIANTS=$'MPH_X=0\nMPH_KEEP_HEAP=0\nMPH_HOST_THREADS=4\nMPH_HOST_THREADS=4 MPH_KEEP_HEAP=0' bash tools/gpu_iter.sh
for i in 1 2 3 4; do tail -1 gpurun_out/r2ac_bench_$i.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); e=j['e2e']; print('$i e2e %.2f ms' % e['ms_per_step'], e['stages_ms'])"; done
