python tools/stage_probe.py 100 2>&1 | grep -A2 "rep 2" 
echo "== hex v16"; python tools/run_once.py hex 128 gls 2 2>&1 | tail -1 | cut -c1-100
echo "== hex v20"; NPB_GLS_VARIANT=20 python tools/run_once.py hex 128 gls 2 2>&1 | tail -1 | cut -c1-100
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -1 gpurun_out/bench_c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c4.json"))
print({k:d[k] for k in ("metric","value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
