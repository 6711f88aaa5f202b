"""Data-parallel training step: one process per GPU, the batch sharded by rows, ONE exchange step — a SUM
all-reduce of the gradients over NCCL (NVLink 5 / NVSwitch) — issued in the order gradients become final
(`linear.*` right after the fused CE backward, so it overlaps the whole BPTT; LSTM weights after the reverse
recurrence; `embed.weight` last) on NCCL's own stream while the compute stream keeps going.

This replaces the reference's `nn.DataParallel(model, device_ids=range(num_gpu))` (train.py:43-44), which is
single-process and cannot work for this model (a Python-list `lengths` is not scattered).
Parity target: the single-process step on the GLOBAL batch — each rank scales its mean loss by
N_rank / N_global, so the summed gradients equal the global-batch gradients.  The encoder head's BatchNorm
uses per-shard batch statistics (as torch DDP / DataParallel replicas would); decoder-only steps are exact.
Greedy decode shards the batch with no communication at all.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import ops


def strided_shard(n_rows, world, rank):
    """Row i of the length-sorted global batch goes to rank i % world: every shard stays sorted descending and
    token counts stay balanced (contiguous chunks would give rank 0 all the long captions)."""
    return np.arange(rank, n_rows, world)


def shard_batch(batch, world, rank):
    """batch: dict(captions[B,T], lengths list, features/pooled [B,*]) of the global, length-sorted batch ->
    this rank's shard plus n_tokens_global."""
    idx = strided_shard(len(batch["lengths"]), world, rank)
    out = {"lengths": [int(batch["lengths"][i]) for i in idx], "n_tokens_global": int(sum(batch["lengths"]))}
    for k, v in batch.items():
        if k not in ("lengths",) and hasattr(v, "__getitem__") and not isinstance(v, (list, int)):
            out[k] = v[idx]
    return out


class GradAllReducer:
    """Callback for ops.decoder_loss(grad_ready=...): starts an asynchronous all-reduce for each group of
    gradients as soon as it has been enqueued, and blocks the compute stream on them only at the end of
    backward.  Works with any backend (`nccl` on GPUs; `gloo` in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.pending = []
        self.order = []     # names in the order they were reduced (introspection / tests)
        self.bytes = 0

    def __call__(self, names, tensors):
        if names is None:
            self.finish()
            return
        self._reduce_group(tensors)
        for n, t in zip(names, tensors):
            self.order.append(n)
            self.bytes += t.numel() * t.element_size()

    def _reduce_group(self, tensors):
        """One asynchronous SUM all-reduce per readiness group.  On NCCL the group's tensors are coalesced into a
        single launch (ncclGroupStart/End): one kernel instead of one per tensor next to the BPTT kernels."""
        tensors = list(tensors)
        if len(tensors) > 1 and dist.get_backend(self.group) == "nccl" and hasattr(dist, "_coalescing_manager"):
            with dist._coalescing_manager(group=self.group, device=tensors[0].device, async_ops=True) as cm:
                for t in tensors:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.pending.append(cm)
        else:
            for t in tensors:
                self.pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for w in self.pending:
            w.wait()
        self.pending = []


class DataParallelStep:
    """head -> decoder loss -> backward (+ overlapped gradient all-reduce) -> clip_gradient + Adam, i.e.
    train.py:137-146 for the models.py pair, per rank.  world == 1 needs no process group."""

    def __init__(self, encoder, decoder, group=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_clip=0.1,
                 optimizer=True):
        self.encoder, self.decoder = encoder, decoder
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.group = group
        self.reducer = GradAllReducer(group) if self.world > 1 else None
        decoder.grad_ready = self.reducer
        self.params = [p for m in (encoder, decoder) if m is not None for p in m.parameters() if p.requires_grad]
        self.lr, self.betas, self.eps, self.grad_clip = lr, betas, eps, grad_clip
        self.optimizer = optimizer
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0
        self._inflight = []   # events of the last steps (world > 1): bounds how far the host runs ahead

    def step(self, inputs, captions, lengths, targets, n_tokens_global=None):
        """inputs: pooled[B,2048] when an encoder is attached, else features[B,E].  Returns this rank's share
        of the global mean loss (sum over ranks = global-batch loss)."""
        if self.world > 1 and self.params and self.params[0].is_cuda:
            # Gradients are consumed on NCCL's stream, so the caching allocator can recycle their blocks only once
            # that work has completed.  A host that runs many steps ahead keeps asking for fresh blocks (cudaMalloc
            # synchronises the device); two steps of run-ahead hide all launch latency and keep the pool steady.
            if len(self._inflight) >= 2:
                self._inflight.pop(0).synchronize()
        for p in self.params:
            p.grad = None
        n_local = int(sum(lengths))
        scale = 1.0 if (self.world == 1 or n_tokens_global is None) else n_local / float(n_tokens_global)
        feats = self.encoder.forward_pooled(inputs) if self.encoder is not None else inputs
        loss = self.decoder.loss(feats, captions, lengths, targets, grad_scale=scale)
        loss.backward()
        if self.world > 1 and self.encoder is not None:   # head gradients: final only after the decoder's backward
            hg = [p.grad for p in self.encoder.parameters() if p.requires_grad and p.grad is not None]
            self.reducer._reduce_group(hg)
            self.reducer.finish()
        if self.optimizer:
            self.t += 1
            live = [(p.data, p.grad.contiguous(), m, v) for p, m, v in zip(self.params, self.m, self.v)
                    if p.grad is not None]
            ops.clamp_adam_multi_([x[0] for x in live], [x[1] for x in live], [x[2] for x in live],
                                  [x[3] for x in live], self.t, self.lr, self.betas, self.eps, self.grad_clip)
        if self.world > 1 and self.params and self.params[0].is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
        return loss.detach()
