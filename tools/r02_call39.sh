set -u
# 2 GPUs, HEAD: parity of the fused exchange against the single-process global-batch step, then a short weak-scaling line
O=gpurun_out/r02m; mkdir -p $O
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_parity.py > $O/parity2.log 2> $O/parity2.err; echo "parity rc=$?"; grep '^{' $O/parity2.log | tail -1 | cut -c1-1200; tail -3 $O/parity2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 20 --warmup 5 --stages --no-greedy --no-extras > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"
python - <<PY
import json
try:
    lines=[l for l in open('$O/bench_2gpu.json') if l.startswith('{')]
    d=json.loads(lines[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'loss',d.get('loss'))
    print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
except Exception as e:
    print('bench2 failed', e); print(open('$O/bench_2gpu.err').read()[-1500:])
PY
