set -u
O=gpurun_out/r02
mkdir -p $O
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tests/multi_gpu_parity.py > $O/multi_gpu_parity3.out 2> $O/multi_gpu_parity3.err; echo "parity rc=$?"; grep '^{' $O/multi_gpu_parity3.out | tail -1 | cut -c1-900; tail -3 $O/multi_gpu_parity3.err | cut -c1-300
for n in $N; do
for v in "X=1" "SNT_DP_PIPELINE=0"; do
env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 30 --warmup 5 --no-greedy --no-extras > $O/bench_${n}gpu_f.json 2> $O/bench_${n}gpu_f.err; echo "bench$n [$v] rc=$?"; tail -2 $O/bench_${n}gpu_f.err | cut -c1-300
python - <<PY
import json
lines=[l for l in open('$O/bench_${n}gpu_f.json') if l.startswith('{')]
d=json.loads(lines[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))
PY
done
done
