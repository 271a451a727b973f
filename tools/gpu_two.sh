# two GPUs of one box: the weak-scaling bench line at N = 2 (one shard per rank, no collective on the data path)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 6 --no-cpu-baseline > gpurun_out/r2_bench_exome_2gpu.json 2> gpurun_out/r2_bench_exome_2gpu.err; echo n2 rc=$?
tail -1 gpurun_out/r2_bench_exome_2gpu.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); e=j['e2e']; print('N=2 value %.3g step %.3f ms e2e %.3g (%.2f ms)' % (j['value'], j['ms_per_step'], e['value'], e['ms_per_step']), e['stages_ms'])"
nproc
