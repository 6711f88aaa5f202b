set -u
O=gpurun_out/r02
mkdir -p $O
nvidia-smi -L | wc -l
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 --stages > $O/bench_${N}gpu_b.json 2> $O/bench_${N}gpu_b.err; echo "bench$N rc=$?"; grep "stages" $O/bench_${N}gpu_b.err | cut -c1-60 | head -12
python - <<PY
import json
d=json.load(open('$O/bench_${N}gpu_b.json'))
print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'strong',d.get('strong_8192'))
print('greedy',d['greedy']['value'])
PY
done
