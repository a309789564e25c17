python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "2d" 2>&1 | tail -30
