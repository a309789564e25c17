python tools/run_once.py tet 69 idw,ls > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_idw|k_ls|k_emit_rows|k_esuel|k_faces|k_sort_rows|k_fill_rows|k_count_nodes" -c 12 -o gpurun_out/prof_k1_idw_ls python tools/run_once.py tet 69 idw,ls > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log
