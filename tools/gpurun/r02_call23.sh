# round 2, call 23 (1 GPU): the full GPU suite, smoke(), the default bench line and the reference arm, on the final build
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench1.json 2> gpurun_out/r02_final_bench1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_final_bench1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err
echo "ref rc=$?"; tail -c 600 gpurun_out/r02_final_bench_ref.json
