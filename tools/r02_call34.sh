set -u
O=gpurun_out/r02m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_fullsize.py tests/test_gpu_trainer.py -x -q > $O/t1.log 2>&1; echo "pytest rc=$?"; tail -2 $O/t1.log
for v in default nosplit; do
  case $v in default) E="SNT_PREP_DEBUG=1";; nosplit) E="SNT_PREP_DEBUG=1 SNT_NO_SPLIT_PROJECTION=1";; esac
  echo "== $v"; env $E timeout 200 python tools/e2e_probe.py 24 > $O/$v.log 2>&1
  grep "prep dbg" $O/$v.log | sed -n '2,5p'; grep -v "per-call (ms)\|prep dbg" $O/$v.log | sed -n '2p'
done
B="python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference --stages"
for v in default nosplit default2 nosplit2; do
  case $v in default*) E="";; nosplit*) E="SNT_NO_SPLIT_PROJECTION=1";; esac
  env $E timeout 300 $B > $O/bench_$v.json 2> $O/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_$v.json') if l.startswith('{')][-1])
    print('$v', 'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'loss', d.get('loss'))
    print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
except Exception as e:
    print('$v failed', e); print(open('$O/bench_$v.err').read()[-1500:])
PY
done
