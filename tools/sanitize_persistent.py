"""One small train step + greedy decode that go through the persistent cooperative recurrence kernels
(`lstm_fwd_persistent_kernel`, `lstm_bwd_persistent_kernel`: spin-waited cross-CTA counters, fence.proxy.async) and the
tcgen05 GEMMs, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_persistent.py
    compute-sanitizer --tool racecheck python tools/sanitize_persistent.py
    compute-sanitizer --tool synccheck python tools/sanitize_persistent.py

Sizes are small (the sanitizer slows kernels 10-100x) but hit the persistent path: H in {256, 512}, more than one row
block, ragged lengths.  Prints the loss and a gradient checksum; the tool's own summary line is the result.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import show_and_tell_b200 as snt


def main():
    H = int(os.environ.get("SNT_SAN_H", "256"))
    B = int(os.environ.get("SNT_SAN_B", "200"))
    E, V = 64, 1000
    torch.manual_seed(0)
    dec = snt.DecoderRNN(E, H, V, 1, precision="bf16").cuda().train()
    b = snt.synthetic.make_batch(B, V, embed=E, seed=3)
    feats = torch.from_numpy(b["features"]).cuda()
    caps = torch.from_numpy(b["captions"]).cuda()
    tg = torch.from_numpy(snt.synthetic.pack_host(b["captions"], b["lengths"])).cuda()
    for it in range(2):
        dec.zero_grad(set_to_none=True)
        loss = dec.loss(feats, caps, b["lengths"], tg)
        loss.backward()
        torch.cuda.synchronize()
        chk = sum(float(p.grad.double().abs().sum()) for p in dec.parameters())
        print(f"step {it}: loss {float(loss):.6f} grad-abs-sum {chk:.6e}", flush=True)
    ids = dec.eval().sample(feats[:64], precision="bf16")
    torch.cuda.synchronize()
    print("greedy ids checksum", int(ids.sum()), flush=True)


if __name__ == "__main__":
    main()
