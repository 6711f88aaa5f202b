"""BASELINE.json configs[3] (embed 512, hidden 1024, 2-layer LSTM, 32k vocab, batch 2048): per-stage GPU time of one
bf16 train step of the decoder (eager, CUDA events around every C-ABI call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import show_and_tell_b200 as snt
from show_and_tell_b200 import parallel
B, E, H, V, L = 2048, 512, 1024, 32000, 2
torch.manual_seed(0)
dec = snt.DecoderRNN(E, H, V, L, precision="bf16").cuda().train()
st = parallel.DataParallelStep(None, dec)
b = snt.synthetic.make_batch(B, V, embed=E, seed=1)
tg = torch.from_numpy(snt.synthetic.pack_host(b["captions"], b["lengths"])).cuda()
feats, caps = torch.from_numpy(b["features"]).cuda(), torch.from_numpy(b["captions"]).cuda()
for _ in range(3): st.step(feats, caps, b["lengths"], tg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): st.step(feats, caps, b["lengths"], tg)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
n = sum(b["lengths"])
flops = 6.0 * n * (4 * H * (E + H) + 4 * H * (H + H) + H * V)
print(f"scaled config: {ms:.3f} ms/step, {B / ms * 1e3:.0f} captions/s, {flops / ms / 1e9:.0f} TFLOP/s algorithmic ({n} tokens)")
snt._lib.profile_begin()
for _ in range(5): st.step(feats, caps, b["lengths"], tg)
for k, (c, t) in sorted(snt._lib.profile_end().items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:24s} {c / 5:4.1f} calls/step {t / 5 * 1e3:9.1f} us/step")
