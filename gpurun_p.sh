python tools/run_once.py tet 40 gls 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gls_mf -s 2 -c 1 -o gpurun_out/prof_gls_mf_h python tools/run_once.py tet 40 gls > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/plain.log | cut -c1-200
