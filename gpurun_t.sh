timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
echo "== tet"; python tools/run_once.py tet 100 gls 2 2>&1 | tail -1 | cut -c1-120
echo "== hex"; python tools/run_once.py hex 128 gls 2 2>&1 | tail -1 | cut -c1-120
echo "== mixed";  python tools/run_once.py mixed 60 gls 2 2>&1 | tail -1 | cut -c1-220
