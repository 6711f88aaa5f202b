set -u
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_trim.py -x -q -m gpu > $O/gputest_multi2.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/gputest_multi2.log
for v in "X=1" "NCCL_MAX_CTAS=32"; do
env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 --stages --no-greedy --no-extras > $O/bench_2gpu_c.json 2> $O/bench_2gpu_c.err; echo "bench2 [$v] rc=$?"; grep "stages" $O/bench_2gpu_c.err | cut -c1-60 | head -3
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r02/bench_2gpu_c.json') if l.startswith('{')]
d=json.loads(lines[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))
PY
done
python -c "
import json
lines=[l for l in open('gpurun_out/r02/bench_2gpu_c.json') if l.startswith('{')]
print(lines[-1][:300])"
