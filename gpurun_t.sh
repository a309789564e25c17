timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python tools/run_once.py tet 100 gls 2 2>&1 | tail -1 | cut -c1-120; }
run NPB_GLS_VARIANT=12
run NPB_GLS_FCAP=4:1272
echo "== hex"; python tools/run_once.py hex 128 gls 2 2>&1 | tail -1 | cut -c1-120
echo "== mixed"; NPB_GLS_FCAP=4:1272 python tools/run_once.py mixed 60 gls 2 2>&1 | tail -1 | cut -c1-220
