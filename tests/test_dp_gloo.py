"""World-size-2 check of the data-parallel host logic on CPU (gloo): strided sharding of the length-sorted batch,
N_rank/N_global loss weighting, the initial broadcast from rank 0, the three contiguous gradient buckets of the flat
buffer all-reduced in readiness order and the staged optimizer (show_and_tell_b200.parallel.DataParallelStep) reproduce
the single-process step - gradients AND updated parameters - on the GLOBAL batch.  The per-rank compute here is the
CPU baseline port (test infrastructure) behind the engine interface; the CUDA executor behind the same class is covered
by tests/test_gpu_step.py and tests/multi_gpu_parity.py (NCCL, 2 GPUs)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, E, H, V, L = 22, 16, 24, 97, 2
LR, CLIP = 1e-2, 0.1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_batch(seed=3):
    import show_and_tell_b200 as snt
    return snt.synthetic.make_batch(B, V, embed=E, seed=seed)


def _model(seed=11):
    from oracle import torch_port as TP
    torch.manual_seed(seed)
    dec = TP.CaptionDecoderCPU(E, H, V, L)
    dec.num_layers = L
    return dec


class PortEngine:
    """The engine interface of parallel.DataParallelStep on the CPU port: prepare / run(phases) / adam / loss / flat.
    Gradients are computed with torch autograd in the first phase and copied into the flat buffer bucket by bucket, in
    the phase the CUDA executor finalises them."""

    def __init__(self, dec):
        from show_and_tell_b200 import engine
        self.dec = dec
        self.flat = engine.FlatParams(None, dec)
        self.loss = torch.zeros(())
        self.scale = 1.0
        self.phases = []

    def prepare(self, inputs, captions, lengths, targets, grad_scale):
        self.batch = (inputs, captions, [int(l) for l in lengths], targets)
        self.scale = grad_scale
        return int(sum(self.batch[2]))

    def set_grad_scale(self, s):
        self.scale = s

    def run(self, phases):
        from show_and_tell_b200 import engine as E_
        self.phases.append(phases)
        f = self.flat
        if phases & E_.PH_FWD:
            inputs, captions, lengths, targets = self.batch
            if targets is None:
                targets = torch.nn.utils.rnn.pack_padded_sequence(captions, lengths, batch_first=True)[0]
            loss = torch.nn.functional.cross_entropy(self.dec(inputs, captions, lengths), targets) * self.scale
            self.autograd = dict(zip(f.names, torch.autograd.grad(loss, f.params)))
            self.loss = loss.detach()
        for bit, bucket in ((E_.PH_BWD_CE, "early"), (E_.PH_BWD_LSTM, "mid"), (E_.PH_BWD_TAIL, "late")):
            if phases & bit:
                lo, hi = f.bucket_range[bucket]
                for n, o, p in zip(f.names, f.offsets, f.params):
                    if lo <= o < hi:
                        f.g[o:o + p.numel()].copy_(self.autograd[n].reshape(-1))

    def adam(self, lo, hi, step, lr, betas, eps, grad_clip):
        f = self.flat
        g = f.g[lo:hi].clamp(-grad_clip, grad_clip)                       # train.py:88-91
        f.m[lo:hi].mul_(betas[0]).add_(g, alpha=1 - betas[0])              # torch.optim.Adam
        f.v[lo:hi].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        f.p[lo:hi].addcdiv_(f.m[lo:hi], (f.v[lo:hi].sqrt() / bc2 ** 0.5).add_(eps), value=-lr / bc1)


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import engine, parallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    dec = _model(seed=11 + 7 * rank)                  # ranks start DIFFERENT: the stepper must broadcast rank 0's
    eng = PortEngine(dec)
    st = parallel.DataParallelStep(None, dec, lr=LR, grad_clip=CLIP, engine=eng)
    ref0 = _model(seed=11)
    for k, v in ref0.state_dict().items():
        assert torch.equal(dec.state_dict()[k], v), k                     # broadcast from rank 0 happened
    losses = []
    for it in range(2):                               # two steps: the second one runs on the UPDATED parameters
        gb = _global_batch(seed=3 + it)
        sh = parallel.shard_batch(gb, world, rank)
        assert sh["lengths"] == sorted(sh["lengths"], reverse=True)       # every shard stays sorted
        loss = st.step(torch.from_numpy(sh["features"]), torch.from_numpy(sh["captions"]), sh["lengths"], None,
                       sh["n_tokens_global"])
        tot = loss.clone()
        dist.all_reduce(tot)
        losses.append(float(tot))
        if it == 0:
            grads0 = {n: st.flat.grad(n).clone().numpy() for n in st.flat.names}
    assert eng.phases[:3] == [engine.PH_FWD | engine.PH_BWD_CE, engine.PH_BWD_LSTM, engine.PH_BWD_TAIL]
    assert st.reducer.order[:3] == ["early", "mid", "late"] and not st.reducer.pending
    if rank == 0:
        q.put((grads0, {k: v.numpy().copy() for k, v in dec.state_dict().items()}, losses, st.reducer.bytes, st.t))
    dist.barrier()
    st.close()
    dist.destroy_process_group()


def test_dp2_step_matches_global_batch_step():
    import show_and_tell_b200 as snt
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    grads0, state, losses, nbytes, t = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, global batch, torch's own clip + Adam (train.py:137-146)
    dec = _model(seed=11)
    opt = torch.optim.Adam(dec.parameters(), lr=LR)
    for it in range(2):
        gb = _global_batch(seed=3 + it)
        targets = torch.from_numpy(snt.synthetic.pack_host(gb["captions"], gb["lengths"]))
        opt.zero_grad()
        ref = torch.nn.functional.cross_entropy(
            dec(torch.from_numpy(gb["features"]), torch.from_numpy(gb["captions"]), gb["lengths"]), targets)
        ref.backward()
        assert abs(losses[it] - float(ref)) < 2e-6 * abs(float(ref))
        if it == 0:
            for k, p in dec.named_parameters():
                np.testing.assert_allclose(grads0[k], p.grad.numpy(), rtol=2e-5, atol=2e-7, err_msg=k)
        for p in dec.parameters():
            p.grad.clamp_(-CLIP, CLIP)
        opt.step()
    for k, v in dec.state_dict().items():
        np.testing.assert_allclose(state[k], v.numpy(), rtol=1e-4, atol=2e-6, err_msg=k)
    assert t == 2
    padded = sum((p.numel() + 63) // 64 * 64 for p in dec.parameters())
    assert nbytes == 2 * 4 * padded                                     # every parameter all-reduced exactly once per step


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
