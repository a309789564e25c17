# round 2, call 16 (1 GPU): compute-sanitizer memcheck over the small-mesh suite (every kernel family, every switch).
set -x
timeout 600 python tools/sanitize_suite.py quick > gpurun_out/r02_sanitize_plain.log 2>&1
echo "plain rc=$?"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 30 python tools/sanitize_suite.py quick > gpurun_out/r02_sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -5 gpurun_out/r02_sanitize_plain.log
grep -c "Invalid\|Error" gpurun_out/r02_sanitize_memcheck.log
tail -8 gpurun_out/r02_sanitize_memcheck.log
