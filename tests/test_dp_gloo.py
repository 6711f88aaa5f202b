"""World-size-2 check of the data-parallel host logic on CPU (gloo): strided sharding of the length-sorted
batch, N_rank/N_global loss weighting, and the readiness-ordered asynchronous gradient all-reduce
(show_and_tell_b200.parallel) reproduce the single-process step on the GLOBAL batch.  The per-rank compute here
is the CPU baseline port (test infrastructure) — the CUDA kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, E, H, V, L = 22, 16, 24, 97, 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_batch():
    import show_and_tell_b200 as snt
    b = snt.synthetic.make_batch(B, V, embed=E, seed=3)
    return b


def _model():
    from oracle import torch_port as TP
    torch.manual_seed(11)
    return TP.CaptionDecoderCPU(E, H, V, L)


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    gb = _global_batch()
    sh = parallel.shard_batch(gb, world, rank)
    assert sh["lengths"] == sorted(sh["lengths"], reverse=True)          # every shard stays sorted
    targets = snt.synthetic.pack_host(sh["captions"], sh["lengths"])
    dec = _model()
    scale = sum(sh["lengths"]) / sh["n_tokens_global"]
    logits = dec(torch.from_numpy(sh["features"]), torch.from_numpy(sh["captions"]), sh["lengths"])
    loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(targets)) * scale
    loss.backward()
    red = parallel.GradAllReducer()
    groups = [["linear.weight", "linear.bias"],
              ["lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0"], ["embed.weight"]]
    named = dict(dec.named_parameters())
    for g in groups:                       # the order ops._DecoderLoss.backward reports gradients in
        red(g, [named[n].grad for n in g])
    red(None, None)                        # end of backward: wait for the collectives
    # deferred mode (what DataParallelStep uses with its optimizer): the end-of-backward call does not wait, the owner
    # waits for the early groups first, updates their parameters, then for the rest
    t1, t2, t3 = torch.ones(5), torch.ones(7), torch.ones(3)
    red2 = parallel.GradAllReducer()
    red2.defer = True
    red2(["a"], [t1])
    red2(["b", "c"], [t2, t3])
    red2(None, None)
    assert len(red2.groups) == 2 and len(red2.pending) >= 2
    assert red2.wait_groups(1) == ["a"] and float(t1[0]) == world and len(red2.groups) == 1
    red2.finish()
    assert float(t2[0]) == world and float(t3[0]) == world and not red2.pending and not red2.groups
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    if rank == 0:
        q.put(({k: p.grad.numpy().copy() for k, p in named.items()}, float(tot), red.order, red.bytes))
    dist.barrier()
    dist.destroy_process_group()


def test_dp2_matches_global_batch():
    import show_and_tell_b200 as snt
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    grads, loss, order, nbytes = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gb = _global_batch()
    dec = _model()
    targets = snt.synthetic.pack_host(gb["captions"], gb["lengths"])
    ref = torch.nn.functional.cross_entropy(
        dec(torch.from_numpy(gb["features"]), torch.from_numpy(gb["captions"]), gb["lengths"]), torch.from_numpy(targets))
    ref.backward()
    assert abs(loss - float(ref)) < 1e-6 * abs(float(ref))
    for k, p in dec.named_parameters():
        np.testing.assert_allclose(grads[k], p.grad.numpy(), rtol=2e-5, atol=2e-7, err_msg=k)
    assert order[:2] == ["linear.weight", "linear.bias"] and order[-1] == "embed.weight"
    assert nbytes == sum(p.numel() * 4 for p in dec.parameters())


def test_strided_shard_balances_tokens():
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    b = snt.synthetic.make_batch(8192, 10000, seed=1)
    for world in (2, 4, 8):
        toks = []
        seen = []
        for r in range(world):
            sh = parallel.shard_batch(b, world, r)
            assert sh["lengths"] == sorted(sh["lengths"], reverse=True)
            toks.append(sum(sh["lengths"]))
            seen.extend(parallel.strided_shard(8192, world, r).tolist())
        assert sorted(seen) == list(range(8192))                   # a partition of the batch
        assert max(toks) - min(toks) <= 8192 // world              # within one token per row of each other
        assert sum(toks) == sum(b["lengths"])
