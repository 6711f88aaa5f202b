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
        self._symm = None          # (handle of g, handle of p) when the flat buffers live in symmetric multicast memory
        if engine is None:
            head = encoder if (encoder is not None and any(p.requires_grad for p in encoder.parameters())) else None
            alloc = self._symmetric_alloc(decoder) if (self.world > 1 and optimizer) else None
            engine = StepEngine(FlatParams(head, decoder, alloc=alloc))
            if alloc is not None:
                self._symm_rendezvous(engine.flat)
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
            self._stream = torch.cuda.Stream(device=self.flat.p.device, priority=-2)
        # one GPU: background stream (default = lowest priority) for the early share of the optimizer, see _step
        self._bgstream = None
        if self._stream is not None and self.world == 1 and os.environ.get("SNT_EARLY_ADAM", "1") != "0":
            self._bgstream = torch.cuda.Stream(device=self.flat.p.device)
            self._bg_ev, self._bg_done = torch.cuda.Event(), torch.cuda.Event()
        if self.world > 1:
            self.sync_from_rank0()

    # ---- fused exchange: gradients, optimizer and parameter broadcast as one kernel over NVSwitch multicast memory ------
    def _symmetric_alloc(self, decoder):
        """-> allocator of symmetric (peer-mapped) fp32 buffers, or None when this job cannot use them (CPU tensors, no
        multicast support, SNT_DP_FUSED=0): the bucketed NCCL all-reduce below is the fallback, decided identically on
        every rank."""
        dev = next(decoder.parameters()).device
        if dev.type != "cuda" or os.environ.get("SNT_DP_FUSED", "1") == "0":
            return None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            from torch._C._distributed_c10d import _SymmetricMemory
            ok = bool(_SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA, dev.index or 0))
        except Exception:   # noqa: BLE001 - API absent in this torch build
            ok = False
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag) == 0:
            return None

        state = self._symm_state = {"failed": False}

        def alloc(n):   # a failed symmetric allocation degrades to a plain buffer; the ranks agree on the outcome later
            try:
                if not state["failed"]:
                    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
                    t.zero_()
                    return t
            except Exception:   # noqa: BLE001
                state["failed"] = True
            return torch.zeros(n, dtype=torch.float32, device=dev)
        return alloc

    def _symm_rendezvous(self, flat):
        import torch.distributed._symmetric_memory as symm_mem
        # every rank must hold symmetric buffers before anyone enters the (collective) rendezvous; whatever goes wrong on
        # any rank - here or in the rendezvous - sends ALL ranks to the NCCL all-reduce path (the buffers stay valid
        # plain CUDA memory)
        flag = torch.tensor([0 if self._symm_state["failed"] else 1], device=flat.p.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag) == 0:
            return
        hg = hp = None
        try:
            name = (self.group or dist.group.WORLD).group_name
            hg = symm_mem.rendezvous(flat.g, name)
            hp = symm_mem.rendezvous(flat.p, name)
            ok = bool(hg.multicast_ptr) and bool(hp.multicast_ptr)
        except Exception:   # noqa: BLE001
            ok = False
        flag = torch.tensor([1 if ok else 0], device=flat.p.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag) == 0:
            return
        self._symm = (hg, hp)

        def shard(lo, hi):   # this rank's contiguous, 256-float aligned share of flat[lo:hi]
            per = (hi - lo + self.world - 1) // self.world
            per = (per + 255) // 256 * 256
            return (min(lo + self.rank * per, hi), min(lo + (self.rank + 1) * per, hi))
        self._shard = shard(0, flat.numel)
        self._bucket_shard = {b: shard(*flat.bucket_range[b]) for b in flat.BUCKETS}
        # the exchange runs on its own default-priority stream: background work beside the high-priority step stream
        # (priority between the step's stream and the executor's side streams: when thread slots free up on an SM the
        # exchange blocks are placed before more column-sum blocks, but never before the step's own kernels)
        self._xstream = torch.cuda.Stream(device=flat.p.device, priority=-1)
        self._xstream2 = torch.cuda.Stream(device=flat.p.device, priority=-1)   # last bucket, when the first two are late
        self._defer_dw = os.environ.get("SNT_DP_DEFER_DW", "1") != "0"
        sms = torch.cuda.get_device_properties(flat.p.device).multi_processor_count
        self._narrow = 8 * max(8, sms - 128)  # blocks of a bucket exchanged beside the cooperative recurrence kernel
        self._pipelined = os.environ.get("SNT_DP_PIPELINE", "1") != "0"
        self._dbg = [] if os.environ.get("SNT_DP_DEBUG") else None     # (bucket, _, events) per exchange: exchange_report()

    def _exchange_bucket(self, bucket, channel, narrow, xstream=None):
        """Exchange + update of one readiness bucket on the exchange stream, ordered after everything enqueued so far on
        the step's stream: barrier (this bucket's gradients are final on every rank) -> snt_dp_adam_shard."""
        hg, hp = self._symm
        cur = torch.cuda.current_stream(self.flat.p.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        lo, hi = self._bucket_shard[bucket]
        dbg = self._dbg
        xstream = xstream or self._xstream
        with torch.cuda.stream(xstream):
            xstream.wait_event(ev)
            if dbg is not None:
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
            hg.barrier(channel)
            if dbg is not None:
                e[1].record()
            self.engine.dp_adam_shard(hg.multicast_ptr, hp.multicast_ptr, lo, hi, self.t, self.lr, self.betas, self.eps,
                                      self.grad_clip, self._narrow if narrow else 0)
            if dbg is not None:
                e[2].record()
                dbg.append((bucket, None, e))

    def exchange_report(self, last=30):
        """SNT_DP_DEBUG=1: mean microseconds of (wait at the barrier, exchange kernel) per bucket over the last exchanges."""
        if not self._dbg:
            return {}
        torch.cuda.synchronize()
        acc = {}
        for bucket, _, e in self._dbg[-3 * last:]:
            a = acc.setdefault(bucket, [0.0, 0.0, 0])
            a[0] += e[0].elapsed_time(e[1]) * 1e3
            a[1] += e[1].elapsed_time(e[2]) * 1e3
            a[2] += 1
        return {b: {"barrier_us": a[0] / a[2], "kernel_us": a[1] / a[2]} for b, a in acc.items()}

    def _fused_update(self):
        """The whole exchange in one piece (a step that ran with optimizer=False, then update()): barrier (every rank's
        gradients are written) -> snt_dp_adam_shard on this rank's share of every bucket -> barrier (every rank's
        parameters have arrived).  Index ownership is the same as in the pipelined step: by bucket."""
        hg, hp = self._symm
        hg.barrier(0)
        for b in self.flat.BUCKETS:
            lo, hi = self._bucket_shard[b]
            self.engine.dp_adam_shard(hg.multicast_ptr, hp.multicast_ptr, lo, hi, self.t, self.lr, self.betas, self.eps,
                                      self.grad_clip)
        hp.barrier(1)

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

    @property
    def grads_are_local(self):
        """True when the exchange is fused with the optimizer: flat.g then holds this rank's own gradients (their sum over
        ranks exists only inside the update), and the Adam moments are kept by the owner rank of each shard."""
        return self._symm is not None

    def gather_moments(self):
        """-> (m, v) complete on every rank (collective).  With the fused exchange each rank maintains the moments of its own
        shard only; this assembles them, e.g. for a checkpoint or a parity check."""
        m, v = self.flat.m, self.flat.v
        if self._symm is not None:
            own = torch.zeros_like(m, dtype=torch.bool)
            for lo, hi in self._bucket_shard.values():
                own[lo:hi] = True
            m, v = m * own, v * own
            for t in (m, v):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return m, v
        return m.clone(), v.clone()

    def update(self):
        """clip_gradient + Adam (train.py:145-146) on the gradients of the last step; several ranks: including their
        exchange.  Called by step() unless `optimizer` is False."""
        eng, flat = self.engine, self.flat
        self.t += 1
        if self.world > 1 and self._symm is not None:
            self._fused_update()
            return
        lo_e, hi_e = flat.bucket_range["early"]
        lo_m, hi_m = flat.bucket_range["mid"]
        lo_l, hi_l = flat.bucket_range["late"]
        if self.world > 1 and self.reducer.pending:
            # update what has landed while the last bucket is still on the wire
            self.reducer.wait(2)
            eng.adam(lo_e, hi_m, self.t, self.lr, self.betas, self.eps, self.grad_clip)
            self.reducer.wait()
            eng.adam(lo_l, hi_l, self.t, self.lr, self.betas, self.eps, self.grad_clip)
        else:
            eng.adam(lo_e, hi_l, self.t, self.lr, self.betas, self.eps, self.grad_clip)

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
        if self._symm is not None and self.optimizer and self._pipelined:
            # fused exchange, bucket by bucket as the gradients become final: linear.weight is reduced, updated and
            # broadcast while the BPTT runs (on the SMs the recurrence leaves free), the LSTM weights during the tail of
            # backward; only embed.weight + head remain for the end of the step
            self.t += 1
            if self._defer_dw and eng.overlaps_dw_out():
                # the executor runs the d_w_out contraction BESIDE the BPTT recurrence (csrc/step.cu): linear.weight is
                # final after BWD_LSTM, so its exchange moves next to the tail of backward and the last bucket goes on a
                # second exchange stream (its own barrier channel), so that it does not queue behind the first two
                eng.run(PH_FWD | PH_BWD_CE | PH_BWD_LSTM)
                self._exchange_bucket("early", 0, narrow=True)
                self._exchange_bucket("mid", 1, narrow=True)
                eng.run(PH_BWD_TAIL)
                self._exchange_bucket("late", 2, narrow=False, xstream=self._xstream2)
                self._xstream.wait_stream(self._xstream2)
            else:
                eng.run(PH_FWD | PH_BWD_CE)
                self._exchange_bucket("early", 0, narrow=True)
                eng.run(PH_BWD_LSTM)
                self._exchange_bucket("mid", 1, narrow=True)
                eng.run(PH_BWD_TAIL)
                self._exchange_bucket("late", 2, narrow=False)
            with torch.cuda.stream(self._xstream):
                self._symm[1].barrier(3)                      # every rank's parameters have arrived
            torch.cuda.current_stream(flat.p.device).wait_stream(self._xstream)
            flat.attach_grads()
            return self._finish_step(cuda)
        if (self.world == 1 and self.optimizer and self._bgstream is not None and not eng.overlaps_dw_out()
                and flat.bucket_range["early"][1] > flat.bucket_range["early"][0]):
            # One GPU: linear.weight - more than half of all parameters - is final after the vocab-CE backward and nothing
            # reads the fp32 master copy again in this step (the backward contraction uses the bf16 copy made in forward),
            # so its share of clip + Adam (HBM-bound, ~30 us) runs as background work beside the BPTT recurrence (HBM idle)
            # instead of after the step, like the fused exchange of several GPUs does.  (Not when the executor runs
            # d_w_out itself beside the BPTT: small batches, engine.overlaps_dw_out.)
            self.t += 1
            eng.run(PH_FWD | PH_BWD_CE)
            cur = torch.cuda.current_stream(flat.p.device)
            self._bg_ev.record(cur)
            with torch.cuda.stream(self._bgstream):
                self._bgstream.wait_event(self._bg_ev)
                eng.adam(*flat.bucket_range["early"], self.t, self.lr, self.betas, self.eps, self.grad_clip)
                self._bg_done.record(self._bgstream)
            eng.run(PH_BWD_LSTM | PH_BWD_TAIL)
            flat.attach_grads()
            cur.wait_event(self._bg_done)
            eng.adam(flat.bucket_range["mid"][0], flat.bucket_range["late"][1], self.t, self.lr, self.betas, self.eps,
                     self.grad_clip)
            return self._finish_step(cuda)
        if self.world == 1 or self._symm is not None:
            eng.run(PH_ALL)      # (fused exchange in one piece: update())
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
            self.update()
        elif self.reducer is not None and self._symm is None:
            self.reducer.wait()
        return self._finish_step(cuda)

    def _finish_step(self, cuda):
        if self.world > 1 and cuda:
            if self.sm_reserve > 0 and self._symm is None:
                _lib.lib().snt_set_sm_reserve(0)
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
        return self.engine.loss
