python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/mcheck.log 2>&1; echo "check rc=$?"; grep -c OK gpurun_out/mcheck.log; grep -v " OK " gpurun_out/mcheck.log | tail -15
df -h /dev/shm | tail -1
