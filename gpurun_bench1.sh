set -x
python bench.py --workload tet40 --steps 3 --warmup 3 --also idw,ls 2> gpurun_out/b40.err | tee gpurun_out/b40.json
tail -5 gpurun_out/b40.err
python bench.py --workload tet69 --steps 2 --warmup 3 --also idw,ls --no-cpu 2> gpurun_out/b69.err | tee gpurun_out/b69.json
tail -5 gpurun_out/b69.err
python bench.py --workload hex128 --steps 2 --warmup 3 --also idw,ls --no-cpu 2> gpurun_out/bh128.err | tee gpurun_out/bh128.json
tail -5 gpurun_out/bh128.err
