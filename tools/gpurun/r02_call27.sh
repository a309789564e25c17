# round 2, call 27 (1 GPU): full GPU suite with the team launch, then the quick bench line
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_final2_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_final2_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r02_final2_bench1.json 2> gpurun_out/r02_final2_bench1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_final2_bench1.json
