set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest_c.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gputest_c.log
B="python bench.py --steps 40 --warmup 5 --stages --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference"
for v in "X=1" "SNT_NO_BIAS_DEFER=1" "SNT_NO_WGRAD_OVERLAP=1" "SNT_NO_BIAS_DEFER=1 SNT_NO_WGRAD_OVERLAP=1"; do
  echo "== $v"
  env $v timeout 300 $B > $O/ab.json 2> $O/ab.err; grep "stages" $O/ab.err | cut -c1-60; python -c "
import json; d=json.load(open('$O/ab.json')); print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])"
done
timeout 300 python tools/one_step.py --greedy > $O/one_greedy_plain2.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|argmax_finish' -s 30 -c 6 -f -o $O/greedy_full python tools/one_step.py --greedy > $O/ncu_greedy_full.log 2>&1
echo "greedy full rc=$?"; tail -2 $O/ncu_greedy_full.log
ncu -i $O/greedy_full.ncu-rep --page raw --csv > $O/greedy_full_raw.csv 2>/dev/null
ncu -i $O/greedy_full.ncu-rep --page source --csv > $O/greedy_full_source.csv 2>/dev/null
ls -la $O | tail -6
