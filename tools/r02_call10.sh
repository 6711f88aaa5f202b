set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest_e.log 2>&1; echo "pytest rc=$?"; tail -3 $O/gputest_e.log
B="python bench.py --steps 40 --warmup 5 --stages --no-cpu-baseline --no-greedy --no-extras --no-gpu-reference"
for v in "X=1" "SNT_STEP_PRIORITY=0" "SNT_NO_BIAS_DEFER=1" "SNT_NO_BACKGROUND=1"; do
  echo "== $v"
  env $v timeout 300 $B > $O/ab.json 2> $O/ab.err; grep "stages" $O/ab.err | cut -c1-60 | head -4; python -c "
import json; d=json.load(open('$O/ab.json')); print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])"
done
