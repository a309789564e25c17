timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_host_paths.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu --also idw,ls > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -1 gpurun_out/bench_c4.err | cut -c1-150
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c4.json"))
print({k:round(d[k],1) for k in ("value","ms_per_step")}, round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
PY
