mkdir -p gpurun_out
nproc > gpurun_out/r2w_nproc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 6 --no-cpu-baseline > gpurun_out/r2w_bench_n8.json 2> gpurun_out/r2w_bench_n8.err; echo n8 rc=$?
timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --e2e-steps 6 --no-cpu-baseline > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo n1 rc=$?
for n in 8 1; do tail -1 gpurun_out/r2w_bench_n$n.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); e=j['e2e']; print('N=$n value %.3g step %.3f ms e2e %.3g (%.2f ms)' % (j['value'], j['ms_per_step'], e['value'], e['ms_per_step']), e['stages_ms'])"; done
