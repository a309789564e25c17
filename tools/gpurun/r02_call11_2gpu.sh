# round 2, call 11 (2 GPUs): final validation - the whole GPU suite (incl. the 2-GPU pytest), smoke, bench at N=1 and N=2
set -x
( time python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r02_gputest11.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_gputest11.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err | cut -c1-250
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n2.err | cut -c1-250
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_ref.json 2> gpurun_out/r02_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n1","r02_bench_n2"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["metric"], "value %.4g ms %.3f kernel %.3f e2e %.4g (%.1f ms)"%(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("   k1", {k: round(v,2) for k,v in d["load_mesh"]["breakdown_ms"].items()}, "wall", round(d["load_mesh"]["wall_s"],2))
    for k in ("device_compute_only","device_gather_after_kernels","k4","e2e_gather_all"):
        if k in d: print("   ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in d[k].items() if a not in ("note","api")})
    for m,v in d.get("also",{}).items(): print("   also", m, "value %.4g ms %.3f kernel %.3f frac %.3f by-step %.3f"%(v["value"], v["ms_per_step"], v["kernel_ms"], v["roofline"]["frac"], v["roofline_by_step_time"]["frac"]))
    for k,v in d.get("configs",{}).items():
        print("   ", k, v.get("error") or {m:("%.4g"%x["value"], "%.3f ms"%x["ms_per_step"], "frac %.3f"%x["roofline"]["frac"], "e2e %.4g"%x["e2e"]["value"]) for m,x in v.get("methods",{}).items()})
    print("   cpu_baseline", d.get("cpu_baseline",{}).get("value"))
print(open("gpurun_out/r02_ref.json").read()[:300])
PY
