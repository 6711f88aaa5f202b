set -u
O=gpurun_out/r02
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tests/multi_gpu_parity.py > $O/multi_gpu_parity2.out 2> $O/multi_gpu_parity2.err; echo "parity rc=$?"; grep '^{' $O/multi_gpu_parity2.out | tail -1; tail -5 $O/multi_gpu_parity2.err | cut -c1-300
for v in "X=1" "SNT_DP_FUSED=0"; do
env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 30 --warmup 5 --stages --no-greedy --no-extras > $O/bench_2gpu_d.json 2> $O/bench_2gpu_d.err; echo "bench2 [$v] rc=$?"; grep "stages" $O/bench_2gpu_d.err | cut -c1-60 | head -12; tail -3 $O/bench_2gpu_d.err | cut -c1-300
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r02/bench_2gpu_d.json') if l.startswith('{')]
d=json.loads(lines[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), d['config'].get('exchange','')[:60])
PY
done
