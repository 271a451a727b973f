mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_device or pipelined" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_exome_2gpu.json 2> gpurun_out/r1e_bench_exome_2gpu.err; echo rc=$?
tail -c 2200 gpurun_out/r1e_bench_exome_2gpu.json; tail -3 gpurun_out/r1e_bench_exome_2gpu.err
