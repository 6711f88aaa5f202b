#!/usr/bin/env bash
# One gpurun call that answers "do the experimental switches work, and what do they buy?" (DESIGN.md §8).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/switch_sweep.sh'
# Results land in gpurun_out/sweep_*.{json,err}; the last lines printed are a summary table.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --stages --no-cpu-baseline --no-gpu-reference --no-extras --no-greedy"
SNT_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_experimental.py -q > gpurun_out/sweep_tests.log 2>&1
echo "experimental tests rc=$? ($(tail -1 gpurun_out/sweep_tests.log))"
# the whole GEMM / parity suites once more with the multicast contraction switched on everywhere it applies
SNT_GEMM_MC=2 timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/sweep_mc_suite.log 2>&1
echo "suite under SNT_GEMM_MC=2 rc=$? ($(tail -1 gpurun_out/sweep_mc_suite.log))"
run() {  # name, then VAR=1 ...
  local name=$1; shift
  env "$@" timeout 120 $B > gpurun_out/sweep_$name.json 2> gpurun_out/sweep_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/sweep_{name}.json").read().strip().splitlines()[-1])
    st = {e["stage"]: e["us_per_step"] for e in d["stages"]}
    print(f"{name:18s} {d['ms_per_step']*1e3:8.1f} us/step  ce_fwd {st.get('snt_vocab_ce_fwd', 0):6.1f}  ce_bwd {st.get('snt_vocab_ce_bwd', 0):6.1f}  "
          f"embed_bwd {st.get('snt_embed_pack_bwd', st.get('snt_embed_pack_bwd_planned', 0)):6.1f}  "
          f"head_bwd {st.get('snt_head_bwd', 0):6.1f}  loss {d['loss']:.6f}")
except Exception as e:
    print(f"{name:18s} FAILED: {e!r}")
PY
}
run baseline SNT_NOOP=1
run lazy_onehot SNT_CEBWD_LAZY=1
run tail_overlap SNT_TAIL_OVERLAP=1
run plan_early SNT_EMB_PLAN_EARLY=1
run multicast2 SNT_GEMM_MC=2
run multicast4 SNT_GEMM_MC=4
run all_four SNT_CEBWD_LAZY=1 SNT_TAIL_OVERLAP=1 SNT_EMB_PLAN_EARLY=1 SNT_GEMM_MC=2
# greedy decode (configs[2]) with and without the multicast vocabulary contraction
for mc in 0 2 4; do
  SNT_GEMM_MC=$mc timeout 120 python - <<'PY'
import os, time, torch
import show_and_tell_b200 as snt
torch.manual_seed(0)
dec = snt.DecoderRNN(256, 512, 10000, 1).cuda().eval()
f = torch.randn(4096, 256, device="cuda")
dec.sample(f, precision="bf16"); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): dec.sample(f, precision="bf16")
e1.record(); torch.cuda.synchronize()
print(f"greedy bf16 SNT_GEMM_MC={os.environ.get('SNT_GEMM_MC')}: {4096 * 20 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e6:.1f} M tokens/s")
PY
done
timeout 240 python tools/gemm_step_shapes.py 2>&1 | tee gpurun_out/sweep_gemm_shapes.txt
