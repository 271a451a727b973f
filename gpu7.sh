python -m pytest tests/test_gpu_peptides.py -x -q 2>&1 | tail -8
