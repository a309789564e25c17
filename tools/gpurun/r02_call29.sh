# round 2, call 29 (1 GPU): ncu --set full of the GLS team launch (Kuhn n = 100, un-chunked), after a plain run of the same command
set -x
RUN_ONCE_CHUNKS=1 python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_team_plain.log 2>&1 &&
RUN_ONCE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:k_gls_mf_team -c 1 -o gpurun_out/r02_prof_gls_team -f python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_team_ncu.log 2>&1
echo "capture rc=$?"
tail -1 gpurun_out/r02_prof_team_plain.log | cut -c1-200
