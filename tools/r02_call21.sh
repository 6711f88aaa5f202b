set -u
O=gpurun_out/r02; mkdir -p $O
n=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tests/multi_gpu_parity.py > $O/multi_gpu_parity3.out 2> $O/multi_gpu_parity3.err; echo "parity rc=$?"; grep '^{' $O/multi_gpu_parity3.out | tail -1 | cut -c1-200
SNT_DP_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 30 --warmup 5 --no-greedy --no-extras > $O/bench_${n}gpu_g.json 2> $O/bench_${n}gpu_g.err; echo "bench$n rc=$?"; grep "\[dp\]" $O/bench_${n}gpu_g.err | head -3 | cut -c1-400
python - <<PY
import json
lines=[l for l in open('$O/bench_${n}gpu_g.json') if l.startswith('{')]
d=json.loads(lines[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))
PY
