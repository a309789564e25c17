# round 2, call 5: full GPU suite, then the tile-shape sweep and a short default bench
set -x
( time python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r02_gputest5.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_gputest5.log
python tools/tile_sweep.py tet203 hex200 > gpurun_out/r02_tile_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02_tile_sweep.log | tail -24
python bench.py --steps 3 --warmup 3 --no-cpu --configs C2_tet69,C5_mixed170 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_bench4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench4.json"))
print(d["metric"], "value %.4g ms %.3f kernel %.3f e2e %.4g (%.1f ms)"%(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print("   k1", {k: round(v,2) for k,v in d["load_mesh"]["breakdown_ms"].items()}, "wall", round(d["load_mesh"]["wall_s"],2))
for k,v in d.get("configs",{}).items():
    print(k, {m:(round(x["value"]), round(x["e2e"]["value"])) for m,x in v.get("methods",{}).items()}, v.get("error"))
PY
