timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/run_once.py mixed 8 idw,ls,gls > gpurun_out/memcheck.log 2>&1; echo "exit $?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck.log | head -20
tail -5 gpurun_out/memcheck.log
