set -x
python bench.py --steps 2 --warmup 3 --no-cpu --also idw,ls > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu --also idw,ls > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
python tools/run_once.py tet 203 gls > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gls_mf -s 2 -c 1 -o gpurun_out/prof_gls_c4 python tools/run_once.py tet 203 gls > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/plain_c4.log | cut -c1-300; tail -2 gpurun_out/ncu_c4.log; wc -l gpurun_out/launches_c4.csv
