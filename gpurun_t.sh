python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload hex200 --steps 3 --warmup 3 --method ls --also idw --no-cpu 2> gpurun_out/bh.err > gpurun_out/bh.json; tail -1 gpurun_out/bh.err
python bench.py --workload tet203 --steps 3 --warmup 3 --method ls --also idw --no-cpu 2> gpurun_out/bt.err > gpurun_out/bt.json; tail -1 gpurun_out/bt.err
python - <<'PY'
import json
for f in ("bh","bt"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "LS nodes/s %.3g ms %.3f kernel_ms %.3f frac %.3f | IDW %.3g ms %.3f frac %.3f | e2e %.3g" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["also"]["idw"]["value"], d["also"]["idw"]["ms_per_step"], d["also"]["idw"]["roofline"]["frac"], d["e2e"]["value"]))
PY
