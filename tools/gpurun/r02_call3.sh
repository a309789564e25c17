# round 2, call 3: pipelined TMA tile kernels (parity + A/B timing), GLS accuracy probe, e2e after the pool warm-up fix
set -x
( time python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_at_size.py --deselect tests/test_gpu_full_size.py -k "not 2d_gls" ) > gpurun_out/r02_gputest3.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_gputest3.log
python tools/gls_accuracy_probe.py big > gpurun_out/r02_gls_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r02_gls_probe.log | tail -80
python bench.py --steps 3 --warmup 3 --no-configs --no-cpu > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"
NPB_TILE_NO_PIPE=1 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu --method idw --also ls > gpurun_out/r02_bench2_nopipe.json 2> gpurun_out/r02_bench2_nopipe.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-configs --no-cpu --workload hex200 --method idw --also ls > gpurun_out/r02_bench2_hex.json 2> gpurun_out/r02_bench2_hex.err; echo "bench rc=$?"
NPB_TILE_NO_PIPE=1 python bench.py --steps 3 --warmup 3 --no-configs --no-cpu --workload hex200 --method idw --also ls > gpurun_out/r02_bench2_hex_nopipe.json 2> gpurun_out/r02_bench2_hex_nopipe.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r02_bench2","r02_bench2_nopipe","r02_bench2_hex","r02_bench2_hex_nopipe"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d["metric"], "value %.4g ms %.3f kernel %.3f frac %.3f e2e %.4g (%.1f ms)"%(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    for m,v in d.get("also",{}).items():
        print("   also", m, "value %.4g ms %.3f kernel %.3f frac %.3f"%(v["value"], v["ms_per_step"], v["kernel_ms"], v["roofline"]["frac"]))
PY
