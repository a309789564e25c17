python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --workload tet69 --steps 3 --warmup 3 --method idw --also ls --no-cpu 2> gpurun_out/b69.err > gpurun_out/b69.json; tail -2 gpurun_out/b69.err
python bench.py --workload hex200 --steps 3 --warmup 3 --method idw --also ls --no-cpu 2> gpurun_out/bh.err > gpurun_out/bh.json; tail -2 gpurun_out/bh.err
python - <<'PY'
import json
for f in ("b69","bh"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "IDW nodes/s %.3g ms %.3f kernel_ms %.3f frac %.3f | LS %.3g ms %.3f frac %.3f | e2e %.3g" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["also"]["ls"]["value"], d["also"]["ls"]["ms_per_step"], d["also"]["ls"]["roofline"]["frac"], d["e2e"]["value"]))
PY
