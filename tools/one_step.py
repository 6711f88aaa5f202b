"""One steady-state training step (BASELINE configs[1] or --config scaled) and, with --greedy, one greedy decode
(configs[2]) between cudaProfilerStart/Stop, for the ncu captures committed under profiles/:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/r02/launches_step.csv python tools/one_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02/step_full \
        python tools/one_step.py

Without ncu it prints the step's device time (CUDA events) and the number of kernels launched, so the same command can
be checked to exit 0 before it is profiled.  The step goes through parallel.DataParallelStep -> snt_step_run exactly as
bench.py's timed region does (ragged batch, optimizer on)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import show_and_tell_b200 as snt
from show_and_tell_b200 import _lib, parallel

CONFIGS = {"default": dict(B=1024, E=256, H=512, V=10000, L=1), "scaled": dict(B=2048, E=512, H=1024, V=32000, L=2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="default", choices=sorted(CONFIGS))
    ap.add_argument("--greedy", action="store_true", help="profile one greedy decode of 4096 features instead")
    ap.add_argument("--prec", default="bf16")
    ap.add_argument("--warm", type=int, default=4)
    a = ap.parse_args()
    c = CONFIGS[a.config]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    enc = snt.EncoderCNN(c["E"], backbone=False, precision=a.prec).to(dev).train()
    dec = snt.DecoderRNN(c["E"], c["H"], c["V"], c["L"], precision=a.prec).to(dev).train()
    rt = torch.cuda.cudart()
    if a.greedy:
        feats = torch.randn(4096, c["E"], device=dev)
        dec.eval()
        for _ in range(2):
            dec.sample(feats, precision=a.prec)
        torch.cuda.synchronize()
        n0 = _lib.lib().snt_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rt.cudaProfilerStart()
        e0.record()
        ids = dec.sample(feats, precision=a.prec)
        e1.record()
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
        print(f"greedy decode: {e0.elapsed_time(e1) * 1e3:.1f} us, {_lib.lib().snt_launch_count(0)} launches, "
              f"ids checksum {int(ids.sum())}")
        return
    st = parallel.DataParallelStep(enc, dec)
    bs = [snt.synthetic.make_batch(c["B"], c["V"], embed=c["E"], seed=1 + k, pooled_dim=2048) for k in range(2)]
    dv = [(torch.from_numpy(np.ascontiguousarray(b["pooled"])).to(dev), torch.from_numpy(b["captions"]).to(dev),
           np.asarray(b["lengths"])) for b in bs]
    for i in range(a.warm):
        st.step(*dv[i & 1])
    torch.cuda.synchronize()
    _lib.lib().snt_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rt.cudaProfilerStart()
    e0.record()
    loss = st.step(*dv[a.warm & 1])
    e1.record()
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    print(f"train step [{a.config}]: {e0.elapsed_time(e1) * 1e3:.1f} us, {_lib.lib().snt_launch_count(0)} launches, "
          f"N={int(sum(dv[a.warm & 1][2]))} tokens, loss {float(loss):.5f}")


if __name__ == "__main__":
    main()
