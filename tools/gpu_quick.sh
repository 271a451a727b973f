mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "shards_do_not or multi_device" 2>&1 | tail -3
