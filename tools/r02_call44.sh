set -u
# 1 GPU, configs[3] (E512/H1024/L2/V32000, batch 2048: the per-step LSTM path): the cell-backward kernel as a link of the
# per-step launch chain, against SNT_NO_PDL=1
O=gpurun_out/r02s; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_step.py -x -q > $O/t1.log 2>&1; echo "pytest rc=$?"; tail -2 $O/t1.log
B="python bench.py --gpus 1 --config scaled --steps 10 --warmup 3 --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference --stages"
for v in default nopdl default2 nopdl2; do
  case $v in default*) E="SNT_X=0";; nopdl*) E="SNT_NO_PDL=1";; esac
  env $E timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_$v.json') if l.startswith('{')][-1])
    print('$v', 'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), 'loss', d.get('loss'), 'launches', d.get('gpu_launches_per_step'))
    print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
except Exception as e:
    print('$v failed', e); print(open('$O/bench_$v.err').read()[-1500:])
PY
done
