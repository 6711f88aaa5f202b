"""GPU micro-benchmark of the tcgen05 contraction core through the C ABI (not a test; run by hand under gpurun).
Prints one line per case: shape, operand majors, beta, microseconds, TFLOP/s."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import show_and_tell_b200 as snt

L = snt._lib
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def run(M, N, K, tA, tB, beta, c_bf16=False, reps=20):
    A = torch.randn((K, M) if tA else (M, K), device="cuda").bfloat16()
    B = torch.randn((N, K) if tB else (K, N), device="cuda").bfloat16()
    Cm = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if c_bf16 else torch.float32)
    f = lambda: L.call("snt_gemm_bf16", tA, tB, M, N, K, 1.0, P(A), A.shape[1], P(B), B.shape[1], beta, P(Cm), N,
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
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"M={M:6d} N={N:6d} K={K:6d} tA={tA} tB={tB} beta={beta} bf16out={int(c_bf16)}  {us:9.1f} us  "
          f"{2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for beta in (0.0, 1.0):
        run(10000, 512, 1024, 1, 0, beta)      # dW_out chunk (MN-major both)
        run(10000, 512, 2048, 1, 0, beta)
        run(1024, 2048, 512, 0, 1, beta)       # recurrent step
        run(1024, 10000, 512, 0, 1, beta)      # logits chunk
    run(12851, 10000, 512, 0, 1, 0.0)          # all logits at once (fp32 out)
    run(12851, 10000, 512, 0, 1, 0.0, True)    # bf16 out
    run(12851, 2048, 256, 0, 1, 0.0)           # input projection
    run(2048, 512, 10000, 0, 0, 0.0)           # dHs chunk (B MN-major)
    run(8192, 8192, 8192, 0, 1, 0.0, True)     # big square, bf16 out
    run(8192, 8192, 8192, 0, 1, 0.0)           # big square, fp32 out
