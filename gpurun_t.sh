python tools/run_once.py tet 203 gls > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_gls_mf -s 2 -c 1 -o gpurun_out/prof_gls_c4_v2 python tools/run_once.py tet 203 gls > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/plain_c4.log | cut -c1-200; tail -2 gpurun_out/ncu_c4.log
