# round 2 profiling call (1 GPU): (1) ncu launch list of the bench command, (2) ncu --set full of the dominant kernels.
# Every ncu run follows a plain run of the same command that exited 0.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-configs"
$B > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench_c4.csv $B > gpurun_out/r02_prof_ncu.json 2> gpurun_out/r02_prof_ncu.err
echo "launch list rc=$?"
RUN_ONCE_CHUNKS=1 python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_gls_plain.log 2>&1 &&
RUN_ONCE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:k_gls_mf -c 3 -o gpurun_out/r02_prof_gls python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_gls_ncu.log 2>&1
echo "gls capture rc=$?"
RUN_ONCE_CHUNKS=1 python tools/run_once.py tet 120 idw,ls > gpurun_out/r02_prof_tiles_plain.log 2>&1 &&
RUN_ONCE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:"k_idw_tile|k_ls_tile|k_tile_pipe" -c 2 -o gpurun_out/r02_prof_tiles python tools/run_once.py tet 120 idw,ls > gpurun_out/r02_prof_tiles_ncu.log 2>&1
echo "tiles capture rc=$?"
python tools/run_once.py tet 69 idw > gpurun_out/r02_prof_k1_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_esuel_star|k_sort_rows_smem|k_fill_rows|k_fill_fsup|k_faces|k_count_nodes" -c 8 -o gpurun_out/r02_prof_k1 python tools/run_once.py tet 69 idw > gpurun_out/r02_prof_k1_ncu.log 2>&1
echo "k1 capture rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
