# round 2, call 1: the whole GPU test suite (incl. the new at-size parity tests) + host facts + a DRAM-traffic probe
# of the GLS kernel at two residencies (metrics-only ncu, after the same command exited 0 plainly)
set -x
nproc; free -g | head -2; df -h /dev/shm | tail -1
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_gputest.log
python tools/run_once.py tet 100 gls > gpurun_out/r02_plain12.log 2>&1 &&
NPB_GLS_CTAS_PER_SM=6 python tools/run_once.py tet 100 gls > gpurun_out/r02_plain6.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum --clock-control none -k regex:k_gls_mf --csv --log-file gpurun_out/r02_gls_dram12.csv python tools/run_once.py tet 100 gls > gpurun_out/r02_ncu12.log 2>&1
NPB_GLS_CTAS_PER_SM=6 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum --clock-control none -k regex:k_gls_mf --csv --log-file gpurun_out/r02_gls_dram6.csv python tools/run_once.py tet 100 gls > gpurun_out/r02_ncu6.log 2>&1
tail -2 gpurun_out/r02_plain12.log gpurun_out/r02_plain6.log
grep -h k_gls_mf gpurun_out/r02_gls_dram12.csv | tail -8
grep -h k_gls_mf gpurun_out/r02_gls_dram6.csv | tail -8
