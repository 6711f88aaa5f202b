set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest_b.log 2>&1; echo "pytest rc=$?"; tail -6 $O/gputest_b.log
timeout 900 python bench.py --steps 20 --warmup 5 --stages --no-gpu-reference > $O/bench_b.json 2> $O/bench_b.err; echo "bench rc=$?"; grep stages $O/bench_b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_b.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'launches/step',d['gpu_launches_per_step'])
print('greedy',d['greedy']['value'],'cfg3',d.get('configs3',{}).get('value'),'trainer',d['f_rows']['trainer_loop']['value'])
PY
timeout 300 python tools/one_step.py --greedy > $O/one_greedy_plain2.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'ArgmaxEpi|LstmFwdEpi|argmax_finish' -s 30 -c 6 -f -o $O/greedy_full python tools/one_step.py --greedy > $O/ncu_greedy_full.log 2>&1
echo "greedy full rc=$?"; tail -2 $O/ncu_greedy_full.log
ncu -i $O/greedy_full.ncu-rep --page raw --csv > $O/greedy_full_raw.csv 2>/dev/null
ncu -i $O/greedy_full.ncu-rep --page source --csv > $O/greedy_full_source.csv 2>/dev/null
ls -la $O | tail -8
