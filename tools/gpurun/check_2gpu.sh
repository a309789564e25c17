python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/mcheck.log 2>&1; echo "check rc=$?"; grep -c OK gpurun_out/mcheck.log; grep MISMATCH gpurun_out/mcheck.log | head; tail -3 gpurun_out/mcheck.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --workload tet69 > gpurun_out/m2.json 2> gpurun_out/m2.err; echo "bench rc=$?"; tail -3 gpurun_out/m2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/m2.json"))
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
PY
