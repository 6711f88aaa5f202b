set -u
O=gpurun_out/r02
mkdir -p $O
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 20 --warmup 5 --stages > $O/bench_${N}gpu_c.json 2> $O/bench_${N}gpu_c.err; echo "bench$N rc=$?"
python - <<PY
import json
lines=[l for l in open('$O/bench_${N}gpu_c.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'greedy',round(d['greedy']['value']))
print('  strong',d.get('strong_8192'))
print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
PY
done
