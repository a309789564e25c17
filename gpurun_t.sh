python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/run_once.py tet 120 gls 3 2>&1 | tail -1 | cut -c1-120
python tools/run_once.py tet 69 gls 3 2>&1 | tail -1 | cut -c1-120
