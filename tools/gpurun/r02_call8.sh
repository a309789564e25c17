# round 2, call 8 (1 GPU): full suite on the padded-centroid build, default bench with all configs, then the ncu
# evidence (launch list of the bench command; full captures of the dominant kernels) - each ncu run after a plain run
set -x
( time python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r02_gputest8.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_gputest8.log
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -12 gpurun_out/r02_bench_n1.err | cut -c1-250
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1.json"))
print(d["metric"], "value %.4g ms %.3f kernel %.3f e2e %.4g (%.1f ms)"%(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print("   k1", {k: round(v,2) for k,v in d["load_mesh"]["breakdown_ms"].items()}, "wall", round(d["load_mesh"]["wall_s"],2), d["load_mesh"].get("host_phases_s"))
for m,v in d.get("also",{}).items(): print("   also", m, "value %.4g ms %.3f kernel %.3f frac %.3f by-step %.3f"%(v["value"], v["ms_per_step"], v["kernel_ms"], v["roofline"]["frac"], v["roofline_by_step_time"]["frac"]))
for k,v in d.get("configs",{}).items():
    print(k, v.get("error") or {m:("%.4g"%x["value"], "%.3f ms"%x["ms_per_step"], "frac %.3f"%x["roofline"]["frac"], "e2e %.4g"%x["e2e"]["value"]) for m,x in v.get("methods",{}).items()}, "k1 %.2f ms load %.2fs"%(v.get("k1_device_ms",0), v.get("load_mesh_wall_s",0)))
print("cpu_baseline", d.get("cpu_baseline",{}).get("value"))
PY
bash tools/gpurun/r02_profile.sh
