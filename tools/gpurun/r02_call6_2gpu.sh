# round 2, call 6 (2 GPUs): multi-GPU parity check (all gather modes x chunk modes, incl. a 117k-node mesh), the 2-GPU
# pytest, then the default bench at N=2
set -x
nvidia-smi -L
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py --big ) > gpurun_out/r02_mgpu_check_2.log 2>&1; echo "check rc=$?"
grep -c " OK " gpurun_out/r02_mgpu_check_2.log; grep -c MISMATCH gpurun_out/r02_mgpu_check_2.log; tail -4 gpurun_out/r02_mgpu_check_2.log | cut -c1-250
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --configs C5_mixed170,C1_tet7 ) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"
tail -6 gpurun_out/r02_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n2.json"))
print("N=2 value %.4g (%.2f ms) compute-only %.4g e2e %.4g (%.1f ms) e2e_all %s"%(d["value"], d["ms_per_step"], d["device_compute_only"]["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("e2e_gather_all",{}).get("value")))
print("k4", d.get("k4")); print("load_mesh wall", d["load_mesh"]["wall_s"])
for m,v in d.get("also",{}).items(): print("also", m, "%.4g"%v["value"], "ms %.3f kernel %.3f by-step frac %.3f"%(v["ms_per_step"], v["kernel_ms"], v["roofline_by_step_time"]["frac"]), v.get("with_nccl_gather"))
for k,v in d.get("configs",{}).items(): print(k, {m:(round(x["value"]), round(x["e2e"]["value"])) for m,x in v.get("methods",{}).items()}, v.get("error"))
PY
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "2d_gls or variants or esuel or pipelined" 2>&1 | tail -4
python tools/tile_sweep.py tet203 hex200 > gpurun_out/r02_tile_sweep2.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02_tile_sweep2.log | tail -40
