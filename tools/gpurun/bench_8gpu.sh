nvidia-smi -L | wc -l; free -g | head -2 | tail -1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 8 --steps 3 --warmup 3 --also idw ) 2> gpurun_out/m8.err | tee gpurun_out/m8.json | cut -c1-300
grep -E "bench\]|real|Error|error" gpurun_out/m8.err | tail -8
