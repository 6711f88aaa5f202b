set -u
mkdir -p gpurun_out/r02
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02/gputest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/r02/gputest_multi.log
P=29611
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P tests/multi_gpu_parity.py > gpurun_out/r02/multi_gpu_parity.out 2> gpurun_out/r02/multi_gpu_parity.err; echo "parity rc=$?"; grep '^{' gpurun_out/r02/multi_gpu_parity.out | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 --stages > gpurun_out/r02/bench_2gpu_a.json 2> gpurun_out/r02/bench_2gpu_a.err; echo "bench2 rc=$?"; tail -30 gpurun_out/r02/bench_2gpu_a.err; cat gpurun_out/r02/bench_2gpu_a.json
