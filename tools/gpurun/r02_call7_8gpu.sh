# round 2, call 7 (8 GPUs): 8-rank parity check on the >100k-node mesh, then the default bench at N=8 (C4 + all configs)
set -x
nvidia-smi -L | wc -l
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py --big ) > gpurun_out/r02_mgpu_check_8.log 2>&1; echo "check rc=$?"
grep -c " OK " gpurun_out/r02_mgpu_check_8.log; grep -c MISMATCH gpurun_out/r02_mgpu_check_8.log; tail -3 gpurun_out/r02_mgpu_check_8.log | cut -c1-250
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"
tail -8 gpurun_out/r02_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n8.json"))
print("N=8 value %.4g (%.2f ms) compute-only %.4g e2e %.4g (%.1f ms) e2e_all %s"%(d["value"], d["ms_per_step"], d["device_compute_only"]["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("e2e_gather_all",{}).get("value")))
print("k4", d.get("k4")); print("load_mesh wall", d["load_mesh"]["wall_s"])
for m,v in d.get("also",{}).items(): print("also", m, "%.4g"%v["value"], "ms %.3f kernel %.3f by-step frac %.3f"%(v["ms_per_step"], v["kernel_ms"], v["roofline_by_step_time"]["frac"]), v.get("with_nccl_gather"))
for k,v in d.get("configs",{}).items(): print(k, {m:(round(x["value"]), round(x["e2e"]["value"])) for m,x in v.get("methods",{}).items()}, v.get("error"))
PY
