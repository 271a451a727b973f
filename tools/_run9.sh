mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" > gpurun_out/r2_tests_multi_device.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests_multi_device.log; tail -3 gpurun_out/r2_tests_multi_device.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 6 --no-cpu-baseline > gpurun_out/r2_bench_exome_8gpu.json 2> gpurun_out/r2_bench_exome_8gpu.err; echo n8 rc=$?
tail -1 gpurun_out/r2_bench_exome_8gpu.json | python -c "
import json,sys
j=json.loads(sys.stdin.read()); e=j['e2e']; print('N=8 value %.3g step %.3f ms e2e %.3g (%.2f ms)' % (j['value'], j['ms_per_step'], e['value'], e['ms_per_step']), e['stages_ms'], 'h2d', e['h2d_bytes_per_step'])"
