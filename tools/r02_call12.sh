set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest_f.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gputest_f.log
timeout 900 python bench.py --steps 20 --warmup 5 --stages --no-gpu-reference > $O/bench_f.json 2> $O/bench_f.err; echo "bench rc=$?"; grep stages $O/bench_f.err | cut -c1-70; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_f.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'launches/step',d['gpu_launches_per_step'])
g=d['greedy']; print('greedy',g['value'],g['ms_per_decode'],'e2e',g['e2e']['value'],'fp32',g['fp32_faithful']['value'])
c3=d['configs3']; print('cfg3',c3['value'],c3['ms_per_step']); 
for s in c3['stages']: print('   ',s['stage'],round(s['us_per_step'],1),round(s['frac'],3))
print('trainer',d['f_rows']['trainer_loop']['value'], 'trim', d['f_rows']['caption_trim'])
print('fp32 train', d.get('fp32_faithful_train'))
PY
timeout 300 python tools/one_step.py --greedy > $O/one_greedy_plain3.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_greedy2.csv python tools/one_step.py --greedy > $O/ncu_greedy2.log 2>&1
echo "greedy launch list rc=$?"; tail -1 $O/one_greedy_plain3.log
