set -u
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02/gputest_a.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02/gputest_a.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_a.json 2> gpurun_out/r02/bench_a.err; echo "bench rc=$?"; tail -5 gpurun_out/r02/bench_a.err; cat gpurun_out/r02/bench_a.json
