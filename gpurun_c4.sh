python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --also idw,ls 2> gpurun_out/c4.err > gpurun_out/c4.json
tail -2 gpurun_out/c4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/c4.json"))
print("GLS %.4g nodes/s ms %.1f fp64frac %.3f e2e %.4g (%.0f ms)" % (d["value"], d["ms_per_step"], d["roofline"]["fp64"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
for m,v in d["also"].items(): print(m, "%.4g nodes/s ms %.2f k2 %.2f frac %.3f" % (v["value"], v["ms_per_step"], v["k2_ms"], v["roofline"]["frac"]))
print(d["load_mesh"]["breakdown_ms"], "wall", d["load_mesh"]["wall_s"], "cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
PY
