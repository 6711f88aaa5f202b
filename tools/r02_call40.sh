set -u
# 2 GPUs: device-side barriers (multimem.red + local poll, folded into the exchange kernel) against torch's symmetric-memory barrier
O=gpurun_out/r02n; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_parity.py > $O/parity2.log 2> $O/parity2.err; echo "parity rc=$?"; grep '^{' $O/parity2.log | tail -1 | cut -c1-1200; tail -3 $O/parity2.err | cut -c1-300
for v in own torch own2 torch2; do
  case $v in own*) E="SNT_DP_OWN_BARRIER=1";; torch*) E="SNT_DP_OWN_BARRIER=0";; esac
  env $E SNT_DP_DEBUG=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 30 --warmup 5 --stages --no-greedy --no-extras > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
  grep "\[dp\] rank 0" $O/bench_$v.err | head -2 | cut -c1-400
  python - <<PY
import json
try:
    lines=[l for l in open('$O/bench_$v.json') if l.startswith('{')]
    d=json.loads(lines[-1]); print('$v N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'loss',d.get('loss'))
    print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
except Exception as e:
    print('$v failed', e); print(open('$O/bench_$v.err').read()[-1500:])
PY
done
