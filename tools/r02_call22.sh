set -u
O=gpurun_out/r02; mkdir -p $O
for n in 8 4; do
SNT_DP_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2965$n bench.py --gpus $n --steps 20 --warmup 5 --stages --no-greedy > $O/bench_${n}gpu_h.json 2> $O/bench_${n}gpu_h.err; echo "bench$n rc=$?"; grep "\[dp\] rank 0" $O/bench_${n}gpu_h.err | head -3 | cut -c1-400
python - <<PY
import json
lines=[l for l in open('$O/bench_${n}gpu_h.json') if l.startswith('{')]
d=json.loads(lines[-1]); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))
print('  strong',d.get('strong_8192'))
print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
PY
done
