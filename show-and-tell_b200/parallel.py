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

Launch the ranks with TORCH_NCCL_HIGH_PRIORITY=1, as bench.py does (the all-reduce CTAs are then placed first when SMs
free up).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops


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
        self.groups = []    # (names, number of handles in self.pending) per readiness group of the current step
        self.bytes = 0
        self.defer = False  # True: the end-of-backward call does not wait; the owner waits group by group
        self.coalesce = True
        self.sm_reserve = int(os.environ.get("SNT_SM_RESERVE", "16"))
        self.reserved = False

    def __call__(self, names, tensors):
        if names is None:
            if not self.defer:
                self.finish()
            return
        if not self.groups and tensors and tensors[0].is_cuda and self.sm_reserve > 0:
            # first collective of this backward: from here on an all-reduce kernel may be running next to the BPTT /
            # embedding / head kernels, so the persistent GEMM grids leave it some SMs (see snt_set_sm_reserve)
            _lib.lib().snt_set_sm_reserve(self.sm_reserve)
            self.reserved = True
        n0 = len(self.pending)
        self._reduce_group(tensors)
        self.groups.append((list(names), len(self.pending) - n0))
        for n, t in zip(names, tensors):
            self.order.append(n)
            self.bytes += t.numel() * t.element_size()

    def wait_groups(self, count):
        """Make the current stream wait for the first `count` still-pending readiness groups; returns their names."""
        names = []
        for _ in range(min(count, len(self.groups))):
            g, k = self.groups.pop(0)
            for w in self.pending[:k]:
                w.wait()
            self.pending = self.pending[k:]
            names += g
        return names

    def _reduce_group(self, tensors):
        """One asynchronous SUM all-reduce per readiness group.  On NCCL the group's tensors are coalesced into a
        single launch (ncclGroupStart/End): one kernel instead of one per tensor next to the BPTT kernels."""
        tensors = list(tensors)
        if (self.coalesce and len(tensors) > 1 and dist.get_backend(self.group) == "nccl"
                and hasattr(dist, "_coalescing_manager")):
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
        self.groups = []
        if self.reserved:
            _lib.lib().snt_set_sm_reserve(0)
            self.reserved = False


class DataParallelStep:
    """head -> decoder loss -> backward (+ overlapped gradient all-reduce) -> clip_gradient + Adam, i.e.
    train.py:137-146 for the models.py pair, per rank.  world == 1 needs no process group."""

    def __init__(self, encoder, decoder, group=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_clip=0.1,
                 optimizer=True, cuda_graph=False, graph_after=3):
        """cuda_graph=True: once the same (input tensors, lengths) signature has been stepped `graph_after` times
        eagerly, forward + backward (+ gradient all-reduce) are captured into a CUDA graph and replayed from then on:
        one graph launch instead of ~60 kernel launches and ~20 C-ABI calls.  The host needs ~1.0 ms to enqueue an eager
        step (measured), which is what bounds a multi-GPU step; the optimizer stays outside the graph (its
        bias-correction scalars change every step).  Signatures that never repeat (real, ragged batches) stay eager.
        Multi-GPU: the collectives are captured too (plain per-tensor all-reduces, all waited for inside the graph); a
        replayed step is one launch, so host-side pauses no longer stall the ranks (2 GPUs: 1.43 ms per step, max 1.54 ms
        over 100 steps, against eager steps with sporadic 3-70 ms stalls).  Call close() before destroying the process
        group."""
        self.encoder, self.decoder = encoder, decoder
        self.cuda_graph, self.graph_after = cuda_graph, graph_after
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # Graph replay with the collectives inside was validated at 2 and 4 GPUs in round 1; the 8-GPU runs of that
            # round used eager launches, so larger jobs stay eager unless SNT_GRAPH_MULTI=1 asks for replay explicitly
            # (SNT_GRAPH_MULTI=0: always eager).
            mode = os.environ.get("SNT_GRAPH_MULTI", "auto")
            if mode == "0" or (mode != "1" and dist.get_world_size(group) > 4):
                self.cuda_graph = False
        self._graphs, self._seen = {}, {}
        self.replayed_kernels = 0   # kernels executed through graph replays (snt_launch_count only sees eager launches)
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
        self.max_run_ahead = 16

    def close(self):
        """Release the captured graphs.  Call before torch.distributed.destroy_process_group(): graphs that contain NCCL
        collectives keep the communicator busy, and barrier()/destroy hang while they are alive (measured, round 1)."""
        self._graphs.clear()
        if self.params and self.params[0].is_cuda:
            torch.cuda.synchronize()

    def _reduce_named(self, grads):
        self.reducer(["encoder.%d" % i for i in range(len(grads))], grads)

    def _fwd_bwd(self, inputs, captions, lengths, targets, n_tokens_global, staged):
        """forward + backward + gradient all-reduce launches.  staged=True (eager, optimizer attached): returns with the
        late groups still in flight; otherwise every collective has been waited for on the current stream."""
        n_local = int(sum(lengths))
        scale = 1.0 if (self.world == 1 or n_tokens_global is None) else n_local / float(n_tokens_global)
        if self.encoder is None:
            feats = inputs
        elif inputs.dim() == 4:                          # images through the frozen cuDNN trunk (models.py:25-29)
            feats = self.encoder(inputs)
        else:
            feats = self.encoder.forward_pooled(inputs)
        if self.reducer is not None:
            self.reducer.defer = staged
        loss = self.decoder.loss(feats, captions, lengths, targets, grad_scale=scale)
        loss.backward()
        if self.world > 1:
            # Readiness order: linear.*, lstm.* (per layer), embed.weight, then the head.  The last two finish only at
            # the very end of backward, so their all-reduce cannot hide behind it; the optimizer therefore first updates
            # the parameters whose gradients are already reduced (70 % of the bytes) while those two are on the wire.
            n_early = max(0, len(self.reducer.groups) - 1) if staged else 0
            if self.encoder is not None:                 # head gradients: final only after the decoder's backward
                hg = [p.grad for p in self.encoder.parameters() if p.requires_grad and p.grad is not None]
                if hg:
                    self._reduce_named(hg)
            if staged:
                self.reducer.wait_groups(n_early)
            else:
                self.reducer.finish()
        return loss

    def _adam(self, late):
        self.t += 1
        live = [(p, p.grad.contiguous(), m, v) for p, m, v in zip(self.params, self.m, self.v) if p.grad is not None]
        first = [x for x in live if id(x[0]) not in late]
        last = [x for x in live if id(x[0]) in late]
        for part in (first, last):
            if part is last and self.world > 1:
                self.reducer.finish()                # embed.weight / head gradients have landed
            if part:
                ops.clamp_adam_multi_([x[0].data for x in part], [x[1] for x in part], [x[2] for x in part],
                                      [x[3] for x in part], self.t, self.lr, self.betas, self.eps, self.grad_clip)

    def _graph_step(self, key, inputs, captions, lengths, targets, n_tokens_global):
        ent = self._graphs.get(key)
        if ent is None:
            g = torch.cuda.CUDAGraph()
            for p in self.params:
                p.grad = None                            # the captured backward allocates the (then static) gradients
            if self.reducer is not None:
                self.reducer.coalesce = False            # plain per-tensor collectives inside the capture
            torch.cuda.synchronize()
            if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
                # the AccumulateGrad nodes were created by the eager steps on the default stream; the capture runs on
                # torch's capture stream: intended, and verified bit-identical (tests/test_gpu_parity.py)
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            n0 = int(_lib.lib().snt_launch_count(0))
            # thread_local: CUDA calls of other threads (e.g. NCCL's watchdog) must not invalidate this capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                loss = self._fwd_bwd(inputs, captions, lengths, targets, n_tokens_global, staged=False)
            kernels = int(_lib.lib().snt_launch_count(0)) - n0   # launches recorded into the graph (not executed yet)
            if self.reducer is not None:
                self.reducer.coalesce = True
            # keep the captured tensors alive; the gradients written by a replay are THIS capture's tensors
            ent = self._graphs[key] = (g, loss, [p.grad for p in self.params], kernels, (inputs, captions, targets))
        ent[0].replay()
        for p, gr in zip(self.params, ent[2]):
            p.grad = gr
        self.replayed_kernels += ent[3]
        return ent[1]

    def step(self, inputs, captions, lengths, targets, n_tokens_global=None):
        """inputs: pooled[B,2048] (or images[B,3,H,W] for an encoder with its backbone) when an encoder is attached,
        else features[B,E].  Returns this rank's share
        of the global mean loss (sum over ranks = global-batch loss)."""
        cuda = bool(self.params) and self.params[0].is_cuda
        if self.world > 1 and cuda:
            # Gradients are consumed on NCCL's stream, so the caching allocator can recycle their blocks only once
            # that work has completed: an unbounded host run-ahead keeps the pool growing.  The bound is generous
            # (16 steps) on purpose: with only 2 steps of queued work every host-side hiccup (measured: sporadic
            # 5-70 ms pauses of the Python thread) stalls the GPUs of ALL ranks through the next all-reduce.
            if len(self._inflight) >= self.max_run_ahead:
                self._inflight.pop(0).synchronize()
        key = None
        if self.cuda_graph and cuda:
            key = (inputs.data_ptr(), captions.data_ptr(), targets.data_ptr(), tuple(int(x) for x in lengths),
                   n_tokens_global, tuple(inputs.shape), tuple(captions.shape))
            self._seen[key] = self._seen.get(key, 0) + 1
            if key not in self._graphs and self._seen[key] <= self.graph_after:
                key = None
        late = set()
        loss = None
        if key is not None:
            try:
                loss = self._graph_step(key, inputs, captions, lengths, targets, n_tokens_global)
            except Exception as e:   # capture refused (driver / allocator state): stay eager from now on, loudly
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({e!r}); continuing with eager launches")
                self.cuda_graph = False
                self._graphs.pop(key, None)
                torch.cuda.synchronize()
                key = None
        if key is None:
            for p in self.params:
                p.grad = None
            if self.reducer is not None:
                self.reducer.coalesce = True
            staged = bool(self.optimizer) and self.world > 1
            loss = self._fwd_bwd(inputs, captions, lengths, targets, n_tokens_global, staged)
            if staged:
                late = {id(p) for n, p in self.decoder.named_parameters() if n == "embed.weight"}
                if self.encoder is not None:
                    late |= {id(p) for p in self.encoder.parameters()}
        if self.optimizer:
            self._adam(late)
        if self.world > 1 and cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
        return loss.detach()
