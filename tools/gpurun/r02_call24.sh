# round 2, call 24 (8 GPUs): the 8-GPU bench line of the final build (GLS chunks from 50 k nodes on: 8 chunks per rank)
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r02_bench8_quick2.json 2> gpurun_out/r02_bench8_quick2.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench8_quick2.json
