( time python -m pytest tests/test_gpu_full_size.py tests/test_gpu_golden.py -m gpu -x -q ) 2>&1 | tail -12
