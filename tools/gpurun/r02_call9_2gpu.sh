# round 2, call 9 (2 GPUs): send/recv gather + device checksum + lazy timers: parity check, bench N=2, GLS variant probe
set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py --big ) > gpurun_out/r02_mgpu_check_2b.log 2>&1; echo "check rc=$?"
grep -c " OK " gpurun_out/r02_mgpu_check_2b.log; grep -c MISMATCH gpurun_out/r02_mgpu_check_2b.log; tail -3 gpurun_out/r02_mgpu_check_2b.log | cut -c1-250
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --configs C2_tet69,C5_mixed170 ) > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; echo "bench rc=$?"
tail -4 gpurun_out/r02_bench_n2b.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n2b.json"))
print("N=2 value %.4g (%.2f ms) compute-only %.4g overlapped %.4g e2e %.4g (%.1f ms) e2e_all %s"%(d["value"], d["ms_per_step"], d["device_compute_only"]["value"], d["device_overlapped_gather"]["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("e2e_gather_all",{}).get("value")))
print("k4", d.get("k4"))
for m,v in d.get("also",{}).items(): print("also", m, "%.4g"%v["value"], "ms %.3f kernel %.3f by-step frac %.3f"%(v["ms_per_step"], v["kernel_ms"], v["roofline_by_step_time"]["frac"]), v.get("with_nccl_gather"))
for k,v in d.get("configs",{}).items(): print(k, {m:(round(x["value"]), round(x["e2e"]["value"])) for m,x in v.get("methods",{}).items()}, v.get("error"))
PY
python tools/gls_variant_probe.py 100 > gpurun_out/r02_gls_variants.log 2>&1; cat gpurun_out/r02_gls_variants.log | tail -9
python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
