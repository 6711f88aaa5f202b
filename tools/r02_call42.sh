set -u
# 1 GPU: chained launches (programmatic dependent launch) as the default, now also head forward / backward and the
# embedding-gradient kernels, gradient buffer zeroed beside the forward pass; against SNT_NO_PDL=1 / SNT_NO_EMB_ZERO_EARLY=1
O=gpurun_out/r02p; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gputest.log
B="python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference --stages"
for v in default nopdl nozero default2 nopdl2; do
  case $v in default*) E="SNT_X=0";; nopdl*) E="SNT_NO_PDL=1";; nozero*) E="SNT_NO_EMB_ZERO_EARLY=1";; esac
  env $E timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_$v.json') if l.startswith('{')][-1])
    print('$v', 'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'loss', d.get('loss'), 'launches', d.get('gpu_launches_per_step'))
    print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
except Exception as e:
    print('$v failed', e); print(open('$O/bench_$v.err').read()[-1500:])
PY
done
