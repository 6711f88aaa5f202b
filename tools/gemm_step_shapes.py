"""The contractions of one training step (E256/H512/V10000/B1024: N = 12666 packed rows, CE chunks of 4736 rows) timed
through snt_gemm_bf16.  Run under gpurun:  python tools/gemm_step_shapes.py   -> one line per shape: us and TFLOP/s."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import show_and_tell_b200 as snt

L = snt._lib
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def time_gemm(M, N, K, tA, tB, c_bf16, reps=20):
    A = torch.randn((K, M) if tA else (M, K), device="cuda").bfloat16()
    B = torch.randn((N, K) if tB else (K, N), device="cuda").bfloat16()
    Cm = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if c_bf16 else torch.float32)
    f = lambda: L.call("snt_gemm_bf16", tA, tB, M, N, K, 1.0, P(A), A.shape[1], P(B), B.shape[1], 0.0, P(Cm), N,
                       1 if c_bf16 else 0, None, L.stream_ptr())
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps, Cm.float().clone()


SHAPES = [  # name, M, N, K, transA, transB, bf16 out
    ("logits-like  hs.W_out^T (all rows)", 12666, 10000, 512, 0, 1, True),
    ("logits-like  one CE chunk", 4736, 10000, 512, 0, 1, True),
    ("dHs = dl.W_out (chunk)", 4736, 512, 10000, 0, 0, False),
    ("dW_out = dl^T.Hs (chunk)", 10000, 512, 4736, 1, 0, False),
    ("Gx' = x.W_ih^T", 12666, 2048, 256, 0, 1, True),
    ("dW_hh = dG^T.Hprev", 2048, 512, 12666, 1, 0, False),
    ("dW_ih = dG^T.X", 2048, 256, 12666, 1, 0, False),
    ("dX = dG.W_ih", 12666, 256, 2048, 0, 0, False),
    ("head  pooled.W_fc^T", 1024, 256, 2048, 0, 1, False),
    ("square 8192^3", 8192, 8192, 8192, 0, 1, True),
]

if __name__ == "__main__":
    for name, M, N, K, tA, tB, bf in SHAPES:
        us0, _ = time_gemm(M, N, K, tA, tB, bf)
        fl = 2.0 * M * N * K
        print(f"{name:38s} M={M:6d} N={N:6d} K={K:6d}  {us0:8.1f} us {fl / us0 / 1e6:7.1f} TF/s", flush=True)
