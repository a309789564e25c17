# round 2, call 28 (2 GPUs): the 2-GPU bench line of the final build (GLS team launch)
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu --no-configs --also "" > gpurun_out/r02_bench2_team.json 2> gpurun_out/r02_bench2_team.err
echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench2_team.json; tail -3 gpurun_out/r02_bench2_team.err
