"""NCCL parity of the data-parallel step (run under torchrun, one rank per GPU; tests/test_gpu_multi.py launches it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multi_gpu_parity.py

Every rank builds DIFFERENT initial weights (the stepper must broadcast rank 0's), takes its strided shard of the same
length-sorted global batches and runs DataParallelStep.step - the exchange fused with the optimizer over NVSwitch
multicast memory (snt_dp_adam_shard; SNT_DP_FUSED=0: three bucketed NCCL all-reduces overlapped with backward, SM
reserve on, staged Adam).  Rank 0 then repeats the steps alone on the GLOBAL batches with the same kernels
(world-size-1 stepper, same initial weights) and compares:
  * step 1: the all-reduced gradients == the global-batch gradients (bf16 mode: 2e-3 rel Frobenius - the shards round
    different partial sums; fp32 mode: 2e-5),
  * after 3 steps: every parameter and Adam moment within the same bars, the summed losses equal the global loss,
  * all ranks hold bit-identical parameters (replicas never diverge).
Exit code 0 = parity green; prints one JSON line with the worst errors."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

E, H, V, L, B_GLOBAL, STEPS = 128, 512, 3000, 2, 600, 3


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(snt, seed, prec, head=True):
    torch.manual_seed(seed)
    enc = snt.EncoderCNN(E, backbone=False, precision=prec).cuda().train() if head else None
    dec = snt.DecoderRNN(E, H, V, L, precision=prec).cuda().train()
    return enc, dec


def main():
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    os.environ.setdefault("NCCL_MAX_CTAS", "16")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    report, ok = {}, True
    for prec, tol in (("fp32", 2e-5), ("bf16", 2e-3)):
        # decoder-only steps are exact (head BatchNorm uses per-shard statistics, a stated deviation): head=False
        enc, dec = build(snt, 100 + rank, prec, head=False)
        st = parallel.DataParallelStep(enc, dec, lr=1e-3, grad_clip=0.1)
        batches = [snt.synthetic.make_batch(B_GLOBAL, V, embed=E, seed=40 + i) for i in range(STEPS)]
        losses, g_first = [], None
        for i, gb in enumerate(batches):
            sh = parallel.shard_batch(gb, world, rank)
            loss = st.step(torch.from_numpy(sh["features"]).cuda(), torch.from_numpy(sh["captions"]).cuda(), sh["lengths"],
                           None, sh["n_tokens_global"])
            tot = loss.clone()
            dist.all_reduce(tot)
            losses.append(float(tot))
            if i == 0:
                g_first = st.flat.g.clone()
                if st.grads_are_local:       # fused exchange: the sum over ranks exists only inside the update kernel
                    dist.all_reduce(g_first)
        # replicas identical?
        mine = st.flat.p.clone()
        ref0 = mine.clone()
        dist.broadcast(ref0, src=0)
        same = torch.equal(mine, ref0)
        flags = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        m_all, v_all = st.gather_moments()          # complete on every rank (owner shards assembled)
        fused = st.grads_are_local
        st.close()
        if rank == 0:
            enc1, dec1 = build(snt, 100, prec, head=False)               # rank 0's initial weights
            st1 = parallel.DataParallelStep(enc1, dec1, lr=1e-3, grad_clip=0.1, distributed=False)
            e = {}
            for i, gb in enumerate(batches):
                l1 = float(st1.step(torch.from_numpy(gb["features"]).cuda(), torch.from_numpy(gb["captions"]).cuda(),
                                    gb["lengths"]))
                e[f"loss{i}"] = abs(losses[i] - l1) / abs(l1)
                if i == 0:
                    e["grad_step1"] = rel(g_first, st1.flat.g)
                    for b in ("early", "mid", "late"):
                        e[f"grad_step1_{b}"] = rel(st.flat.slice(g_first, b), st1.flat.slice(st1.flat.g, b))
            e["params_after"] = rel(st.flat.p, st1.flat.p)
            e["adam_m_after"] = rel(m_all, st1.flat.m)
            e["adam_v_after"] = rel(v_all, st1.flat.v)
            e["fused_exchange"] = bool(fused)
            e["replicas_identical"] = bool(int(flags))
            bad = [k for k, v in e.items() if k not in ("replicas_identical", "fused_exchange")
                   and not (v < (tol if "loss" not in k else 10 * tol))]
            if bad or not e["replicas_identical"]:
                ok = False
                e["FAILED"] = bad
            report[prec] = e
    okt = torch.tensor([int(ok)], device="cuda")
    dist.broadcast(okt, src=0)
    if rank == 0:
        print(json.dumps({"world": world, "ok": ok, "report": report}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(okt) else 1)


if __name__ == "__main__":
    main()
