# round 2, call 12 (1 GPU): validation of the hand-written scan / partition + ramped chunks, then a bench
set -x
( time python -m pytest tests -m gpu -q -x --durations=5 ) > gpurun_out/r02_gputest12.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r02_gputest12.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1b.err | cut -c1-250
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1b.json"))
print(d["metric"], "value %.4g ms %.3f kernel %.3f e2e %.4g (%.1f ms) launches %d"%(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
print("   k1", {k: round(v,2) for k,v in d["load_mesh"]["breakdown_ms"].items()}, "wall", round(d["load_mesh"]["wall_s"],2))
for m,v in d.get("also",{}).items(): print("   also", m, "value %.4g ms %.3f kernel %.3f frac %.3f by-step %.3f"%(v["value"], v["ms_per_step"], v["kernel_ms"], v["roofline"]["frac"], v["roofline_by_step_time"]["frac"]))
for k,v in d.get("configs",{}).items():
    print("   ", k, v.get("error") or {m:("%.4g"%x["value"], "%.3f ms"%x["ms_per_step"], "frac %.3f"%x["roofline"]["frac"], "e2e %.4g"%x["e2e"]["value"]) for m,x in v.get("methods",{}).items()})
PY
