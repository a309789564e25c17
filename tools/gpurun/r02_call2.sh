# round 2, call 2: full GPU suite on the new pipeline (no -x), then the default bench and the reference arm (short)
set -x
( time python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r02_gputest2.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02_gputest2.log
( time python bench.py --steps 3 --warmup 3 --config-steps 2 ) > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/r02_bench1.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02_ref1.json 2> gpurun_out/r02_ref1.err; echo "ref rc=$?"
tail -3 gpurun_out/r02_ref1.err
