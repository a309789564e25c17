# round 2, call 4: full GPU suite (accurate dense GLS reroute, pipelined tiles with aligned rings, K1 row sort in shared
# memory + batched esuel, GLS original rows in shared memory), then A/B timings
set -x
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r02_gputest4.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02_gputest4.log
python bench.py --steps 3 --warmup 3 --no-configs --no-cpu > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench3.err
python bench.py --steps 3 --warmup 3 --no-configs --no-cpu --workload hex200 --method idw --also ls > gpurun_out/r02_bench3_hex.json 2> gpurun_out/r02_bench3_hex.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench3_hex.err
python - <<'PY'
import json
for f in ("r02_bench3","r02_bench3_hex"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d["metric"], "value %.4g ms %.3f kernel %.3f frac %.3f e2e %.4g (%.1f ms)"%(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("   k1", {k: round(v,2) for k,v in d["load_mesh"]["breakdown_ms"].items()}, "wall", round(d["load_mesh"]["wall_s"],2))
    for m,v in d.get("also",{}).items():
        print("   also", m, "value %.4g ms %.3f kernel %.3f frac %.3f"%(v["value"], v["ms_per_step"], v["kernel_ms"], v["roofline"]["frac"]))
PY
