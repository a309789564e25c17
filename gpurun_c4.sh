set -x
free -g | head -2; nproc
( time python bench.py --steps 2 --warmup 3 --also idw,ls ) 2> gpurun_out/c4.err | tee gpurun_out/c4.json
tail -8 gpurun_out/c4.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
