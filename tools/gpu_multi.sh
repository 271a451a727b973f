mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_exome_${N}gpu.json 2> gpurun_out/r1f_bench_exome_${N}gpu.err; echo rc=$?
tail -c 600 gpurun_out/r1f_bench_exome_${N}gpu.err | tail -3
python - <<PY
import json
j=json.loads(open('gpurun_out/r1f_bench_exome_${N}gpu.json').read().strip().split('\n')[-1])
print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step'], j['e2e']['stages_ms'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
