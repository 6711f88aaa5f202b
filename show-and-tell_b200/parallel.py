"""Data-parallel training step: one process per GPU, the batch sharded by rows, ONE exchange step — a SUM
all-reduce of the gradients over NCCL (NVLink 5 / NVSwitch) — issued in the order gradients become final
(`linear.*` right after the vocab-CE backward, so it overlaps the whole BPTT; the LSTM weights after the reverse
recurrence; `embed.weight` and the head last) on NCCL's own stream while the compute stream keeps going.

This replaces the reference's `nn.DataParallel(model, device_ids=range(num_gpu))` (train.py:43-44), which is
single-process and cannot work for this model (a Python-list `lengths` is not scattered).
Parity target: the single-process step on the GLOBAL batch — each rank scales its mean loss by
N_rank / N_global, so the summed gradients equal the global-batch gradients.  The encoder head's BatchNorm
uses per-shard batch statistics (as torch DDP / DataParallel replicas would); decoder-only steps are exact.
Greedy decode shards the batch with no communication at all.

The step itself is the native executor (engine.StepEngine -> snt_step_run): no autograd graph, no per-stage host round
trip.  The host needs ~0.1 ms to enqueue a step, so ragged batches (new lengths every step) run eagerly at GPU speed
and no CUDA-graph capture is involved.  Parameters, gradients and Adam moments live in flat buffers laid out in
readiness order (engine.FlatParams): each of the three gradient buckets is one contiguous all-reduce.

Launch the ranks with TORCH_NCCL_HIGH_PRIORITY=1 and NCCL_MAX_CTAS=16, as bench.py does: the all-reduce CTAs are then
placed first when SMs free up, and an all-reduce in flight never holds more than the 20 SMs the cooperative recurrence
kernels (128 CTAs that must all be resident) leave free - otherwise the BPTT launch waits for the whole collective.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import PH_ALL, PH_BWD_CE, PH_BWD_LSTM, PH_BWD_TAIL, PH_FWD, FlatParams, StepEngine


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


class BucketReducer:
    """Asynchronous SUM all-reduce of contiguous gradient buckets: `start(tensor)` right after the phase that finalises
    the bucket, `wait(k)` before its consumer (the optimizer).  Any backend (`nccl` on GPUs; `gloo` in the CPU
    tests).  The handles keep the bucket tensors alive until waited for."""

    def __init__(self, group=None):
        self.group = group
        self.pending = []
        self.bytes = 0
        self.order = []

    def start(self, name, tensor):
        self.pending.append((name, tensor, dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True)))
        self.order.append(name)
        self.bytes += tensor.numel() * tensor.element_size()

    def wait(self, count=None):
        """Make the current stream (CUDA) / the caller (CPU) wait for the oldest `count` buckets (all when None)."""
        n = len(self.pending) if count is None else min(count, len(self.pending))
        done = []
        for _ in range(n):
            name, _t, work = self.pending.pop(0)
            work.wait()
            done.append(name)
        return done


class DataParallelStep:
    """head -> decoder loss -> backward (+ overlapped gradient all-reduce) -> clip_gradient + Adam, i.e.
    train.py:137-146 for the models.py pair, per rank.  world == 1 needs no process group.

    `engine`: anything with prepare(inputs, captions, lengths, targets, grad_scale) -> n_tokens, run(phases),
    adam(lo, hi, step, lr, betas, eps, grad_clip), `.loss` and `.flat` (a FlatParams); defaults to the CUDA executor.
    The CPU tests inject a torch-based stand-in to exercise this class's bucket logic under gloo."""

    def __init__(self, encoder, decoder, group=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_clip=0.1,
                 optimizer=True, engine=None, sm_reserve=None, distributed=True):
        """distributed=False: a single-process stepper even inside an initialised process group (e.g. the global-batch
        reference of the multi-GPU parity test)."""
        self.encoder, self.decoder = encoder, decoder
        self.group = group
        self.world = dist.get_world_size(group) if (distributed and dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if engine is None:
            head = encoder if (encoder is not None and any(p.requires_grad for p in encoder.parameters())) else None
            engine = StepEngine(FlatParams(head, decoder))
        self.engine = engine
        self.flat = engine.flat
        self.params = self.flat.params
        self.lr, self.betas, self.eps, self.grad_clip = lr, betas, eps, grad_clip
        self.optimizer = optimizer
        self.t = 0
        self.reducer = BucketReducer(group) if self.world > 1 else None
        self.sm_reserve = int(os.environ.get("SNT_SM_RESERVE", "16")) if sm_reserve is None else int(sm_reserve)
        self._inflight = []   # events of the last steps (world > 1): bounds how far the host runs ahead
        self.max_run_ahead = 16
        # The step runs on its own HIGH-priority stream (ordered after the caller's current stream at entry, and the
        # caller's stream after it at exit).  The executor's side streams have default (= lowest) priority, so whatever
        # they carry - the output-bias column sums, the head backward - is background work for the block scheduler: it
        # fills the SMs the cooperative recurrence kernels leave free and never delays them (csrc/step.cu).
        self._stream = None
        if self.flat.p.is_cuda and os.environ.get("SNT_STEP_PRIORITY", "1") != "0":
            self._stream = torch.cuda.Stream(device=self.flat.p.device, priority=-1)
        if self.world > 1:
            self.sync_from_rank0()

    # the reference's DataParallel re-replicates the module from device 0 every step (train.py:43-44); with one process
    # per GPU the replicas are made identical once, here, and stay identical because every rank applies the same
    # all-reduced gradients.  Adam moments and the step count start from zero on every rank.
    def sync_from_rank0(self):
        with torch.no_grad():
            dist.broadcast(self.flat.p, src=0, group=self.group)
            for m in (self.encoder, self.decoder):
                if m is None:
                    continue
                for b in m.buffers():
                    if b.is_floating_point():
                        dist.broadcast(b, src=0, group=self.group)

    def close(self):
        """Wait for outstanding collectives (call before torch.distributed.destroy_process_group())."""
        if self.reducer is not None:
            self.reducer.wait()
        if self.flat.p.is_cuda:
            torch.cuda.synchronize()

    @property
    def m(self):
        return [self.flat.m[o:o + p.numel()].view(p.shape) for o, p in zip(self.flat.offsets, self.params)]

    @property
    def v(self):
        return [self.flat.v[o:o + p.numel()].view(p.shape) for o, p in zip(self.flat.offsets, self.params)]

    def _features_in(self, inputs):
        if inputs.dim() == 4:                                 # images through the frozen cuDNN trunk (models.py:25-29)
            if self.encoder is None or not getattr(self.encoder, "has_backbone", False):
                raise RuntimeError("image inputs need an EncoderCNN with its backbone")
            return self.encoder.pooled(inputs)
        return inputs

    def step(self, inputs, captions, lengths, targets=None, n_tokens_global=None):
        """inputs: pooled[B,2048] (or images[B,3,H,W] for an encoder with its backbone) when an encoder head is
        attached, else features[B,E].  targets=None: pack(captions, lengths) (eval.py:91), gathered on the device.
        Returns this rank's share of the global mean loss (sum over ranks = global-batch loss)."""
        if self._stream is None:
            return self._step(inputs, captions, lengths, targets, n_tokens_global)
        cur = torch.cuda.current_stream(self.flat.p.device)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            loss = self._step(inputs, captions, lengths, targets, n_tokens_global)
        cur.wait_stream(self._stream)
        return loss

    def _step(self, inputs, captions, lengths, targets=None, n_tokens_global=None):
        eng, flat = self.engine, self.flat
        cuda = flat.p.is_cuda
        if not flat.intact():
            raise RuntimeError("a parameter was moved off the flat buffer (module.to()/cuda() after the stepper was "
                               "built): build the DataParallelStep after placing the modules")
        if self.world > 1 and cuda and len(self._inflight) >= self.max_run_ahead:
            self._inflight.pop(0).synchronize()               # bound the host's run-ahead (NCCL holds the buckets)
        inputs = self._features_in(inputs)
        if flat.encoder is not None and flat.encoder.training and flat.encoder.bn.track_running_stats:
            flat.encoder.bn.num_batches_tracked += 1          # what nn.BatchNorm1d.forward does in train()
        n_local = eng.prepare(inputs, captions, lengths, targets, 1.0)
        if self.world > 1 and n_tokens_global is not None:
            eng.set_grad_scale(n_local / float(n_tokens_global))
        if self.world == 1:
            eng.run(PH_ALL)
        else:
            red = self.reducer
            reserve = cuda and self.sm_reserve > 0
            eng.run(PH_FWD | PH_BWD_CE)
            if reserve:                                       # collectives in flight from here on: leave them some SMs
                _lib.lib().snt_set_sm_reserve(self.sm_reserve)
            red.start("early", flat.slice(flat.g, "early"))   # linear.*: overlaps all of BPTT
            eng.run(PH_BWD_LSTM)
            red.start("mid", flat.slice(flat.g, "mid"))       # lstm.*: overlaps the embedding gradient / head backward
            eng.run(PH_BWD_TAIL)
            red.start("late", flat.slice(flat.g, "late"))     # embed.weight + head: final only now
        flat.attach_grads()
        if self.optimizer:
            self.t += 1
            lo_e, hi_e = flat.bucket_range["early"]
            lo_m, hi_m = flat.bucket_range["mid"]
            lo_l, hi_l = flat.bucket_range["late"]
            if self.world > 1:
                # update what has landed while the last bucket is still on the wire
                self.reducer.wait(2)
                eng.adam(lo_e, hi_m, self.t, self.lr, self.betas, self.eps, self.grad_clip)
                self.reducer.wait()
                eng.adam(lo_l, hi_l, self.t, self.lr, self.betas, self.eps, self.grad_clip)
            else:
                eng.adam(lo_e, hi_l, self.t, self.lr, self.betas, self.eps, self.grad_clip)
        elif self.world > 1:
            self.reducer.wait()
        if self.world > 1 and cuda:
            if self.sm_reserve > 0:
                _lib.lib().snt_set_sm_reserve(0)
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
        return eng.loss
