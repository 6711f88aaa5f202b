"""Vocabulary projection + cross-entropy, forward and backward, timed alone through the C ABI with CUDA events:
the recompute path (snt_vocab_ce_fwd / snt_vocab_ce_bwd) against the stored-numerator training path
(snt_vocab_ce_train_fwd / _bwd), and their outputs compared with each other and with a float64 torch evaluation on the
same bf16-rounded operands.  Run under gpurun:

    python tools/ce_probe.py            # cfg2 (N=12666, H=512, V=10000) and cfg3 (N=25702, H=1024, V=32000)
    python tools/ce_probe.py small      # adds small / ragged shapes with the full fp64 check
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import show_and_tell_b200 as snt

L = snt._lib
P = L.ptr


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def run(N, H, V, check64=False, seed=0, wscale=0.1):
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    hs = (torch.rand(N, H, device=dev, generator=g) * 2 - 1).bfloat16()
    w = (torch.rand(V, H, device=dev, generator=g) * 2 - 1) * wscale
    b = torch.randn(V, device=dev, generator=g) * 0.1
    tg = torch.randint(0, V, (N,), device=dev, generator=g)
    p = L.PREC["bf16"]
    lib = L.lib()
    st = L.stream_ptr()
    ws_a = torch.empty(lib.snt_vocab_ce_workspace_bytes(p, N, H, V), dtype=torch.uint8, device=dev)
    ws_b = torch.empty(lib.snt_vocab_ce_train_workspace_bytes(p, N, H, V), dtype=torch.uint8, device=dev)
    out = {}
    for name in ("recompute", "stored"):
        lse, loss = torch.empty(N, device=dev), torch.empty((), device=dev)
        d_hs, d_w, d_b = torch.empty(N, H, device=dev), torch.empty(V, H, device=dev), torch.empty(V, device=dev)
        if name == "recompute":
            fwd = lambda: L.call("snt_vocab_ce_fwd", p, P(hs), P(w), P(b), P(tg), N, H, V, P(lse), P(loss), P(ws_a),
                                 ws_a.numel(), st)
            bwd = lambda: L.call("snt_vocab_ce_bwd", p, P(hs), P(w), P(b), P(tg), P(lse), None, 1.0, N, H, V, P(d_hs),
                                 P(d_w), P(d_b), P(ws_a), ws_a.numel(), st)
        else:
            u = torch.empty(N, (V + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
            inv_s, hs_s = torch.empty(N, device=dev), torch.empty(N, H, device=dev, dtype=torch.bfloat16)
            wb = torch.empty(V, H, device=dev, dtype=torch.bfloat16)
            fwd = lambda: L.call("snt_vocab_ce_train_fwd", p, P(hs), P(w), P(b), P(tg), N, H, V, P(lse), P(loss), P(u),
                                 P(inv_s), P(hs_s), P(wb), P(ws_b), ws_b.numel(), st)
            bwd = lambda: L.call("snt_vocab_ce_train_bwd", p, P(u), P(inv_s), P(hs_s), P(wb), None, 1.0, N, H, V,
                                 P(d_hs), P(d_w), P(d_b), P(ws_b), ws_b.numel(), st)
        t_f = timed(fwd)
        fwd()
        t_b = timed(bwd)
        torch.cuda.synchronize()
        out[name] = dict(t_f=t_f, t_b=t_b, loss=float(loss), lse=lse.clone(), d_hs=d_hs.clone(), d_w=d_w.clone(),
                         d_b=d_b.clone())
    flags = L.read_flags()
    rel = lambda a, r: float((a.double() - r.double()).norm() / r.double().norm().clamp_min(1e-30))
    a, s = out["recompute"], out["stored"]
    fl = 2.0 * N * H * V
    print(f"N={N} H={H} V={V}  flags={flags}")
    for name, o in out.items():
        tot = o["t_f"] + o["t_b"]
        print(f"  {name:10s} fwd {o['t_f']:8.1f} us  bwd {o['t_b']:8.1f} us  total {tot:8.1f} us  "
              f"({3 * fl / tot / 1e6:7.1f} TF/s algorithmic fwd+bwd)  loss {o['loss']:.6f}")
    print(f"  stored vs recompute: lse {rel(s['lse'], a['lse']):.2e}  d_hs {rel(s['d_hs'], a['d_hs']):.2e}  "
          f"d_w {rel(s['d_w'], a['d_w']):.2e}  d_b {rel(s['d_b'], a['d_b']):.2e}")
    if check64:
        hs64, w64 = hs.double(), w.bfloat16().double()
        logits = hs64 @ w64.t() + b.double()
        lse64 = torch.logsumexp(logits, 1)
        pr = torch.softmax(logits, 1)
        pr[torch.arange(N, device=dev), tg] -= 1.0
        pr /= N
        ref = dict(lse=lse64, d_hs=pr @ w64, d_w=pr.t() @ hs64, d_b=pr.sum(0))
        loss64 = float((lse64 - logits[torch.arange(N, device=dev), tg]).mean())
        for name, o in out.items():
            print(f"  {name:10s} vs fp64: loss {abs(o['loss'] - loss64) / abs(loss64):.2e}  " +
                  "  ".join(f"{k} {rel(o[k], ref[k]):.2e}" for k in ("lse", "d_hs", "d_w", "d_b")))
    return out


if __name__ == "__main__":
    if "small" in sys.argv:
        for N, H, V in ((1, 8, 5), (77, 64, 1000), (300, 128, 257), (1000, 256, 2500), (2048, 512, 10000)):
            run(N, H, V, check64=True)
        run(500, 128, 3000, check64=True, wscale=2.0)      # wide logit range: exercises the pilot shift
    run(12666, 512, 10000)
    run(25702, 1024, 32000)
