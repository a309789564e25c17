python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py > gpurun_out/mcheck.log 2>&1
grep -E "rank [01]/|Error|error:|Traceback|File " gpurun_out/mcheck.log | head -60
