# round 2, call 20 (8 GPUs): multi-GPU check incl. the flag-slice scenarios, then the 8-GPU bench line (no CPU arm, no configs)
set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/r02_multi_check_8.log 2>&1
echo "check rc=$?"
grep -c " OK" gpurun_out/r02_multi_check_8.log; grep -c MISMATCH gpurun_out/r02_multi_check_8.log
grep "flag slice" gpurun_out/r02_multi_check_8.log | grep "rank 7" | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r02_bench8_quick.json 2> gpurun_out/r02_bench8_quick.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench8_quick.json
tail -3 gpurun_out/r02_bench8_quick.err
