TAG=r2v TESTK="config or shape or carry or long_peptides" VARIANTS=$'MPH_X=0\nMPH_REPLAY_PER_SM=6\nMPH_SIDE_REPLAY=k1\nMPH_SIDE_REPLAY=0\nMPH_SIDE_REPLAY=0 MPH_MERGE_CTAS=0\nMPH_SIDE_REPLAY=0 MPH_MERGE_CTAS=1184' bash tools/gpu_iter.sh
for w in normal; do timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/r2v_bench_$w.json 2> gpurun_out/r2v_bench_$w.err; tail -1 gpurun_out/r2v_bench_$w.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('normal step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], j['roofline']['kernel_ms'])"; done
