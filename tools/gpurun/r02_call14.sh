# round 2, call 14 (1 GPU): un-chunked ncu --set full captures of the GLS kernel and of the IDW / LS tile kernels.
# Every ncu run follows a plain run of the same command that exited 0.
set -x
RUN_ONCE_CHUNKS=1 python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_gls_plain.log 2>&1 &&
RUN_ONCE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"k_gls_mf<12>" -c 1 -o gpurun_out/r02_prof_gls_full -f python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_gls_ncu.log 2>&1
echo "gls capture rc=$?"
RUN_ONCE_CHUNKS=1 python tools/run_once.py tet 120 idw,ls > gpurun_out/r02_prof_tiles_plain.log 2>&1 &&
RUN_ONCE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"k_idw_tile|k_ls_tile|k_tile_pipe" -c 2 -o gpurun_out/r02_prof_tiles_full -f python tools/run_once.py tet 120 idw,ls > gpurun_out/r02_prof_tiles_ncu.log 2>&1
echo "tiles capture rc=$?"
tail -2 gpurun_out/r02_prof_gls_plain.log gpurun_out/r02_prof_tiles_plain.log
ls -la gpurun_out/*_full.ncu-rep
