set -u
O=gpurun_out/r02j; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_feed.py tests/test_gpu_trainer.py -x -q > $O/t1.log 2>&1; echo "pytest feed+trainer rc=$?"; tail -2 $O/t1.log
# wide batch (8192 captions on one GPU): two row blocks per CTA in the BPTT against one
B="python bench.py --gpus 1 --steps 10 --warmup 3 --scaling strong --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference --stages"
for v in default rb1; do
  case $v in default) E="";; rb1) E="SNT_PERSIST_RB=1";; esac
  env $E timeout 300 $B > $O/strong_$v.json 2> $O/strong_$v.err; echo "strong $v rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/strong_$v.json') if l.startswith('{')][-1])
    print('$v', 'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))
    print('  ', [(s['stage'],round(s['us_per_step'],1)) for s in d['stages']])
except Exception as e:
    print('$v failed', e); print(open('$O/strong_$v.err').read()[-1500:])
PY
done
# source-level capture of the largest kernel of the step (fused vocab-CE forward with stored numerators)
timeout 300 python tools/one_step.py > $O/one_step_plain.log 2>&1 && tail -1 $O/one_step_plain.log &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:CeStoreEpi -c 1 -f -o $O/ce_store python tools/one_step.py > $O/ncu_ce.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_ce.log
ncu -i $O/ce_store.ncu-rep --page raw --csv > $O/ce_store_raw.csv 2>/dev/null
ncu -i $O/ce_store.ncu-rep --page source --csv > $O/ce_store_source.csv 2>/dev/null
ls -la $O | tail -8
