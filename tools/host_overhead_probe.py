"""How long does the HOST take to enqueue one training step (no GPU sync)?  If this approaches the GPU step time the
job is launch-bound.  python tools/host_overhead_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import show_and_tell_b200 as snt
from show_and_tell_b200 import parallel
B, E, H, V = 1024, 256, 512, 10000
torch.manual_seed(0)
enc = snt.EncoderCNN(E, backbone=False, precision="bf16").cuda().train()
dec = snt.DecoderRNN(E, H, V, 1, precision="bf16").cuda().train()
st = parallel.DataParallelStep(enc, dec)
b = snt.synthetic.make_batch(B, V, embed=E, seed=1, pooled_dim=2048)
tg = torch.from_numpy(snt.synthetic.pack_host(b["captions"], b["lengths"])).cuda()
pooled, caps = torch.from_numpy(b["pooled"]).cuda(), torch.from_numpy(b["captions"]).cuda()
for _ in range(10): st.step(pooled, caps, b["lengths"], tg)
torch.cuda.synchronize()
for n in (8, 8, 8):
    t0 = time.perf_counter()
    for _ in range(n): st.step(pooled, caps, b["lengths"], tg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step; until GPU done {1e3 * (t2 - t0) / n:.3f} ms/step")
