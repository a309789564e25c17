# round 2, call 19 (2 GPUs): multi-GPU check incl. the flag-slice scenarios, then the 2-GPU bench line
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/r02_multi_check_2.log 2>&1
echo "check rc=$?"
grep -c " OK" gpurun_out/r02_multi_check_2.log; grep -c MISMATCH gpurun_out/r02_multi_check_2.log
grep "flag slice" gpurun_out/r02_multi_check_2.log | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r02_bench2_quick.json 2> gpurun_out/r02_bench2_quick.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench2_quick.json
tail -5 gpurun_out/r02_bench2_quick.err
