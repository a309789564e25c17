python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "tet 69" "hex 64" "mixed 24"; do python tools/run_once.py $cfg gls 3 2>&1 | tail -1 | cut -c1-260; done
