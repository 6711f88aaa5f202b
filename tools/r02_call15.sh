set -u
O=gpurun_out/r02
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/gputest_h.log 2>&1; echo "pytest rc=$?"; tail -15 $O/gputest_h.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-reference --no-cpu-baseline > $O/bench_h.json 2> $O/bench_h.err; echo "bench rc=$?"; tail -3 $O/bench_h.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02/bench_h.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
g=d['greedy']; print('greedy',g['value'],g['ms_per_decode'],'fp32',g['fp32_faithful'])
print('fp32 train', d.get('fp32_faithful_train'))
print('trim', d['f_rows']['caption_trim']['large'])
PY
SNT_FP32_FFMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
