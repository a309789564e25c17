# ncu evidence for profiles/ (one ncu invocation per gpurun call; each only after the same command exited 0 plainly)
# call 1: launch list of the default bench
#   python bench.py --steps 2 --warmup 3 --no-cpu --also idw,ls > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv \
#       python bench.py --steps 2 --warmup 3 --no-cpu --also idw,ls > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
# call 2: full capture of the interior-node GLS launch (11 GPU-minutes at C4: prefer `tet 100`, 2.5 minutes)
python tools/run_once.py tet 100 gls > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gls_mf -s 2 -c 1 -o gpurun_out/prof_gls python tools/run_once.py tet 100 gls > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200; tail -2 gpurun_out/ncu.log
