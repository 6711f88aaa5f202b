set -u
O=gpurun_out/r02
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_trim.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_step.py -m gpu -x -q > $O/gputest_g.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gputest_g.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline > $O/bench_g.json 2> $O/bench_g.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_g.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
g=d['greedy']; print('greedy',g['value'],g['ms_per_decode'],'e2e',g['e2e']['value'])
c3=d['configs3']; print('cfg3',c3['value'],c3['ms_per_step']); 
for s in c3['stages']: print('   ',s['stage'],round(s['us_per_step'],1),round(s['frac'],3))
print('trainer',d['f_rows']['trainer_loop']['value'], 'trim', d['f_rows']['caption_trim']['large'])
PY
for v in "X=1" "SNT_NO_FUSED_BPTT_STEP=1"; do
env $v timeout 300 python tools/one_step.py --config scaled | tail -1
done
timeout 300 python tools/one_step.py --config scaled > $O/one_step_scaled_plain2.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step_scaled2.csv python tools/one_step.py --config scaled > $O/ncu_step_scaled2.log 2>&1
echo "scaled launch list rc=$?"
