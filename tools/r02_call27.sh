# early weight preparation: why does the host-input loop lose what the resident loop gains?
set -u
O=gpurun_out/r02f2; mkdir -p $O
B="python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference --stages"
for v in default conn32 noprep noprep_conn32; do
  case $v in default) E="";; conn32) E="CUDA_DEVICE_MAX_CONNECTIONS=32";; noprep) E="SNT_NO_EARLY_PREP=1";; noprep_conn32) E="SNT_NO_EARLY_PREP=1 CUDA_DEVICE_MAX_CONNECTIONS=32";; esac
  env $E timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_$v.json') if l.startswith('{')][-1])
    print('$v', 'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],4))
    print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
except Exception as e:
    print('$v failed', e); print(open('$O/bench_$v.err').read()[-1500:])
PY
done
