TAG=r2ad TESTK="config or shape or pipelined or golden or splice or fs" VARIANTS=$'MPH_X=0\nMPH_HOST_THREADS=4' bash tools/gpu_iter.sh
for i in 1 2; do tail -1 gpurun_out/r2ad_bench_$i.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); e=j['e2e']; print('$i e2e %.2f ms' % e['ms_per_step'], e['stages_ms'])"; done
