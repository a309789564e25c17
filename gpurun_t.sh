python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in "tet 40" "hex 64" "mixed 24"; do python tools/run_once.py $cfg gls 2 2>&1 | tail -2; done
python bench.py --workload tet69 --steps 2 --warmup 3 --also idw --no-cpu 2> gpurun_out/b69.err > gpurun_out/b69.json; tail -3 gpurun_out/b69.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/b69.json"))
print("GLS nodes/s", d["value"], "ms", d["ms_per_step"], "fp64 frac", d["roofline"]["fp64"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"])
PY
