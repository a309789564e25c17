python tools/run_once.py tet 120 idw,ls > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_idw_tile|k_ls_tile" -c 2 -o gpurun_out/prof_tiles python tools/run_once.py tet 120 idw,ls > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log
