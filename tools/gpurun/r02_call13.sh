# round 2, call 13 (1 GPU): star esuel with __match_any_sync (parity + timing), then the ncu evidence on the final build
set -x
( time python -m pytest tests -m gpu -q -x -k "not at_size and not full_size" ) > gpurun_out/r02_gputest13.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_gputest13.log
python -m pytest tests/test_gpu_at_size.py -m gpu -q -x -k "c4" 2>&1 | tail -3
bash tools/gpurun/r02_profile.sh
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_prof_plain.json"))
print(d["metric"], "value %.4g ms %.3f e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]))
print("   k1", {k: round(v,2) for k,v in d["load_mesh"]["breakdown_ms"].items()})
PY
