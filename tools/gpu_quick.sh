mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "packer_entry or pipelined" 2>&1 | tail -5
