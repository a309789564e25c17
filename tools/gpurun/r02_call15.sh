# round 2, call 15 (1 GPU): un-chunked ncu --set full capture of the GLS kernels (all three size-class launches of a
# Kuhn n = 100 pass; ncu matches the base name, template arguments are not part of it).
set -x
RUN_ONCE_CHUNKS=1 python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_gls_plain.log 2>&1 &&
RUN_ONCE_CHUNKS=1 ncu --set full --clock-control none --import-source on -k regex:k_gls_mf -c 3 -o gpurun_out/r02_prof_gls_full -f python tools/run_once.py tet 100 gls > gpurun_out/r02_prof_gls_ncu.log 2>&1
echo "gls capture rc=$?"
ls -la gpurun_out/*_full.ncu-rep
