set -u
# 1 GPU: every contraction of the step (and the small kernels between two contractions) launched with programmatic
# stream serialization (SNT_PDL_ALL=1) against plain stream order
O=gpurun_out/r02o; mkdir -p $O
SNT_PDL_ALL=1 timeout 600 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py tests/test_gpu_gemm.py -x -q > $O/t1.log 2>&1; echo "pytest(pdl) rc=$?"; tail -2 $O/t1.log
B="python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference --stages"
for v in default pdl default2 pdl2; do
  case $v in default*) E="SNT_PDL_ALL=0";; pdl*) E="SNT_PDL_ALL=1";; esac
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
