"""cProfile of the host side of the eager training step (which Python/ctypes calls cost the ~1 ms per step)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    st.step(pooled, caps, b["lengths"], tg)
    if _ % 8 == 7: torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
