"""Host side of the native step executor (`snt_step_run`, include/snt_b200.h): the teacher-forced train step of
train.py:137-146 for the models.py pair without the autograd glue of ops.py.

  FlatParams  one contiguous fp32 buffer each for parameters, gradients and the two Adam moments; the modules'
              nn.Parameters become views into it ("fused grad-clamp + Adam over a flat parameter
              buffer, which also makes the all-reduce a single contiguous buffer").  Tensors are laid out in the order
              their gradients become final during backward - linear.* | lstm.* (top layer first) | embed.weight, head -
              so each readiness bucket is one contiguous slice.  Pure torch: also used by the CPU tests.
  StepEngine  fills the C descriptor and calls snt_step_run phase by phase; owns the step workspace.  CUDA only - there
              is no CPU fallback: constructing it without the library or a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

PH_FWD, PH_BWD_CE, PH_BWD_LSTM, PH_BWD_TAIL, PH_ALL = 1, 2, 4, 8, 15
MAX_LAYERS = 8
_ALIGN = 64          # floats: every tensor starts on a 256-byte boundary of the flat buffers


class SntStep(C.Structure):
    """struct snt_step of include/snt_b200.h, field for field."""
    _P8 = C.c_void_p * MAX_LAYERS
    _fields_ = [
        ("struct_bytes", C.c_int32), ("prec", C.c_int32), ("L", C.c_int32), ("T", C.c_int32),
        ("training", C.c_int32), ("reserved", C.c_int32),
        ("B", C.c_int64), ("E", C.c_int64), ("H", C.c_int64), ("V", C.c_int64), ("K", C.c_int64),
        ("cap_stride", C.c_int64),
        ("batch_sizes", C.c_void_p), ("input", C.c_void_p), ("captions", C.c_void_p), ("targets", C.c_void_p),
        ("w_fc", C.c_void_p), ("b_fc", C.c_void_p), ("bn_w", C.c_void_p), ("bn_b", C.c_void_p),
        ("bn_rm", C.c_void_p), ("bn_rv", C.c_void_p),
        ("bn_momentum", C.c_float), ("bn_eps", C.c_float),
        ("w_emb", C.c_void_p), ("w_out", C.c_void_p), ("b_out", C.c_void_p),
        ("w_ih", _P8), ("w_hh", _P8), ("b_ih", _P8), ("b_hh", _P8),
        ("d_w_fc", C.c_void_p), ("d_b_fc", C.c_void_p), ("d_bn_w", C.c_void_p), ("d_bn_b", C.c_void_p),
        ("d_w_emb", C.c_void_p), ("d_w_out", C.c_void_p), ("d_b_out", C.c_void_p),
        ("d_w_ih", _P8), ("d_w_hh", _P8), ("d_b_ih", _P8), ("d_b_hh", _P8),
        ("d_features", C.c_void_p),
        ("grad_scale", C.c_float), ("pad_", C.c_float),
        ("loss", C.c_void_p), ("ws", C.c_void_p), ("ws_bytes", C.c_int64),
    ]


def batch_sizes(lengths, max_steps=None):
    """lengths (list / numpy / CPU tensor of ints sorted descending, data_loader.py:50) -> (int32 batch_sizes[T], N):
    what pack_padded_sequence computes at models.py:51, with its error messages (enforce_sorted=True)."""
    l = np.asarray(lengths, dtype=np.int64)
    if l.ndim != 1 or l.size == 0:
        raise RuntimeError("lengths must be a non-empty 1-D sequence")
    if l.size > 1 and (l[:-1] < l[1:]).any():
        raise RuntimeError("`lengths` array must be sorted in decreasing order")
    if l[-1] < 1:
        raise RuntimeError("Length of all samples has to be greater than 0, but found an element in "
                           "'lengths' that is <= 0")
    T = int(l[0])
    if max_steps is not None and T > max_steps:
        raise RuntimeError(f"max(lengths)={T} exceeds the {max_steps} timesteps available (captions.shape[1] + 1)")
    # batch_sizes[t] = #{i : l_i > t} = B - #{i : l_i <= t}
    bs = (l.size - np.cumsum(np.bincount(l, minlength=T + 1))[:T]).astype(np.int32)
    return bs, int(l.sum())


class FlatParams:
    """Flat fp32 storage for the trainable parameters of (encoder head, decoder), in gradient-readiness order."""

    BUCKETS = ("early", "mid", "late")   # linear.weight | linear.bias, lstm.* | embed.weight + head

    def __init__(self, encoder, decoder, alloc=None):
        """alloc(numel) -> a zeroed fp32 CUDA tensor: used for the parameter and the gradient buffer, so that a
        data-parallel caller can place them in symmetric (peer-mapped, multicast) memory."""
        self.encoder, self.decoder = encoder, decoder
        L = decoder.num_layers
        named = []                                           # (bucket, name, parameter)
        # linear.bias is final one phase after linear.weight: its column sums run beside the BPTT (csrc/step.cu)
        named += [("early", "linear.weight", decoder.linear.weight), ("mid", "linear.bias", decoder.linear.bias)]
        for k in reversed(range(L)):
            for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                named.append(("mid", f"lstm.{n}_l{k}", getattr(decoder.lstm, f"{n}_l{k}")))
        named.append(("late", "embed.weight", decoder.embed.weight))
        if encoder is not None:
            named += [("late", "encoder.fc.weight", encoder.resnet.fc.weight),
                      ("late", "encoder.fc.bias", encoder.resnet.fc.bias),
                      ("late", "encoder.bn.weight", encoder.bn.weight), ("late", "encoder.bn.bias", encoder.bn.bias)]
        named = [(b, n, p) for b, n, p in named if p.requires_grad]
        self.names = [n for _, n, _ in named]
        self.params = [p for _, _, p in named]
        dev = self.params[0].device
        off, self.offsets, self.bucket_range = 0, [], {}
        for b in self.BUCKETS:
            start = off
            for bb, _, p in named:
                if bb == b:
                    self.offsets.append(off)
                    off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            self.bucket_range[b] = (start, off)
        # offsets were appended bucket by bucket, which is the order of `named` already (early, mid, late)
        self.numel = off
        if alloc is None:
            alloc = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)   # noqa: E731
        self.p = alloc(off)
        self.g = alloc(off)
        self.m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.gviews = []
        with torch.no_grad():
            for o, p in zip(self.offsets, self.params):
                n = p.numel()
                self.p[o:o + n].copy_(p.detach().reshape(-1).float())
                p.data = self.p[o:o + n].view(p.shape)       # the module now reads and writes the flat buffer
                self.gviews.append(self.g[o:o + n].view(p.shape))
        self.index = {n: i for i, n in enumerate(self.names)}

    def intact(self):
        """False once a parameter has been moved off the flat buffer (e.g. by module.to(...))."""
        base = self.p.data_ptr()
        return all(p.data_ptr() == base + 4 * o for o, p in zip(self.offsets, self.params))

    def attach_grads(self):
        """param.grad = its slice of the flat gradient buffer (what the step executor writes)."""
        for p, g in zip(self.params, self.gviews):
            if p.grad is not g:
                p.grad = g

    def grad(self, name):
        return self.gviews[self.index[name]]

    def slice(self, buf, bucket):
        lo, hi = self.bucket_range[bucket]
        return buf[lo:hi]


class StepEngine:
    """snt_step_run bound to one (encoder head, decoder) pair and its FlatParams."""

    def __init__(self, flat: FlatParams, precision=None):
        self.flat = flat
        enc, dec = flat.encoder, flat.decoder
        self.dev = flat.p.device
        if self.dev.type != "cuda":
            raise _lib.SntError("StepEngine needs CUDA tensors: there is no CPU fallback")
        self.lib = _lib.lib()
        self._prec_override = precision
        self.prec = precision or dec.precision
        self.L = dec.num_layers
        self.V, self.E = dec.embed.weight.shape
        self.H = dec.lstm.hidden_size
        self.K = enc.resnet.fc.in_features if enc is not None else 0
        self._loss_ring = torch.zeros(256, dtype=torch.float32, device=self.dev)   # one slot per step, reused after 256
        self._loss_i = 0
        self.loss = self._loss_ring[0]
        self.ws = None
        self._keep = None
        d = self.d = SntStep()
        d.struct_bytes = C.sizeof(SntStep)
        d.L, d.E, d.H, d.V, d.K = self.L, self.E, self.H, self.V, self.K
        self.refresh_pointers()

    def refresh_pointers(self):
        """(Re)read every parameter / gradient address (they are constant unless the module is moved)."""
        f, d = self.flat, self.d
        enc, dec = f.encoder, f.decoder
        gp = lambda name: f.grad(name).data_ptr() if name in f.index else None
        d.w_emb, d.w_out, d.b_out = dec.embed.weight.data_ptr(), dec.linear.weight.data_ptr(), dec.linear.bias.data_ptr()
        d.d_w_emb, d.d_w_out, d.d_b_out = gp("embed.weight"), gp("linear.weight"), gp("linear.bias")
        for k in range(self.L):
            for fld, n in (("w_ih", "weight_ih"), ("w_hh", "weight_hh"), ("b_ih", "bias_ih"), ("b_hh", "bias_hh")):
                getattr(d, fld)[k] = getattr(dec.lstm, f"{n}_l{k}").data_ptr()
                getattr(d, "d_" + fld)[k] = gp(f"lstm.{n}_l{k}")
        if enc is not None:
            bn = enc.bn
            d.w_fc, d.b_fc = enc.resnet.fc.weight.data_ptr(), enc.resnet.fc.bias.data_ptr()
            d.bn_w, d.bn_b = bn.weight.data_ptr(), bn.bias.data_ptr()
            d.bn_rm, d.bn_rv = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
            d.bn_momentum, d.bn_eps = float(bn.momentum), float(bn.eps)
            d.d_w_fc, d.d_b_fc = gp("encoder.fc.weight"), gp("encoder.fc.bias")
            d.d_bn_w, d.d_bn_b = gp("encoder.bn.weight"), gp("encoder.bn.bias")

    def _workspace(self, B, n_max):
        nb = self.lib.snt_step_workspace_bytes(_lib.PREC[self.prec], self.L, B, n_max, self.E, self.H, self.V, self.K)
        if nb < 0:
            raise _lib.SntError("snt_step_workspace_bytes: bad sizes")
        if self.ws is None or self.ws.numel() < nb:
            self.ws = None                                   # release before allocating the larger one
            self.ws = torch.empty(int(nb), dtype=torch.uint8, device=self.dev)
        return self.ws

    def prepare(self, inputs, captions, lengths, targets=None, grad_scale=1.0):
        """Fill the descriptor for one batch.  inputs: pooled[B,K] (head attached) or features[B,E]; captions int64
        [B,Tc]; lengths sorted descending; targets int64 [N] or None (gathered on the device: pack(captions, lengths),
        eval.py:91).  -> number of packed rows N."""
        d = self.d
        self.prec = self._prec_override or self.flat.decoder.precision     # the module's mode may change between steps
        if not (inputs.is_cuda and captions.is_cuda):
            raise _lib.SntError("snt_b200 ops need CUDA tensors: there is no CPU fallback")
        if inputs.dtype != torch.float32 or not inputs.is_contiguous():
            inputs = inputs.contiguous().float()
        if captions.dtype != torch.int64:
            captions = captions.long()
        if captions.dim() != 2 or captions.shape[0] != inputs.shape[0]:
            raise RuntimeError("captions must be [B, Tc] with the batch size of the inputs")
        if captions.numel() and captions.stride(1) != 1:
            captions = captions.contiguous()
        B, Tc = captions.shape
        want = self.K if self.K else self.E
        if inputs.dim() != 2 or inputs.shape[1] != want:
            raise RuntimeError(f"inputs must be [B, {want}]")
        bs, n_tok = batch_sizes(lengths, Tc + 1)
        if int(bs[0]) != B:
            raise RuntimeError(f"len(lengths)={int(bs[0])} does not match the batch size {B}")
        T = int(bs.shape[0])
        if targets is None and T > Tc:
            raise RuntimeError("targets=None needs captions at least max(lengths) wide")
        if targets is not None:
            if targets.dtype != torch.int64 or not targets.is_contiguous():
                targets = targets.long().contiguous()
            if targets.numel() != n_tok:
                raise RuntimeError(f"Expected input batch_size ({n_tok}) to match target batch_size ({targets.numel()}).")
        ws = self._workspace(B, B * (Tc + 1))                # sized for the widest batch of this caption width
        d.prec = _lib.PREC[self.prec]
        d.T, d.B = T, B
        d.training = 1 if (self.flat.encoder is not None and self.flat.encoder.training) else 0
        d.cap_stride = captions.stride(0) if captions.numel() else 0
        d.batch_sizes = bs.ctypes.data
        d.input, d.captions = inputs.data_ptr(), captions.data_ptr()
        d.targets = targets.data_ptr() if targets is not None else None
        d.grad_scale = float(grad_scale)
        d.ws, d.ws_bytes = ws.data_ptr(), ws.numel()
        self._loss_i = (self._loss_i + 1) % self._loss_ring.numel()
        self.loss = self._loss_ring[self._loss_i]            # this step's loss (valid until the ring wraps)
        d.loss = self.loss.data_ptr()
        self._keep = (bs, inputs, captions, targets)         # alive until the next prepare()
        return n_tok

    def set_grad_scale(self, scale):
        """Data-parallel: this rank's share N_rank / N_global of the global mean loss (gradients carry it too)."""
        self.d.grad_scale = float(scale)

    def run(self, phases=PH_ALL):
        _lib.call("snt_step_run", C.c_void_p(C.addressof(self.d)), int(phases), _lib.stream_ptr())

    def overlaps_dw_out(self):
        """True when run(PH_BWD_CE | PH_BWD_LSTM [| ...]) of the prepared batch computes d_w_out beside the BPTT recurrence:
        linear.weight's gradient is then final after BWD_LSTM only (include/snt_b200.h: snt_step_overlaps_dw_out)."""
        return bool(self.lib.snt_step_overlaps_dw_out(int(self.d.prec), int(self.d.B), int(self.d.H)))

    def profile(self, enable=True):
        _lib.check(self.lib.snt_step_profile(1 if enable else 0), "snt_step_profile")

    def profile_read(self):
        """-> {stage: ms of the last run}; synchronises the device."""
        ms = (C.c_float * 22)()
        _lib.check(self.lib.snt_step_profile_read(ms, 22), "snt_step_profile_read")
        v = list(ms)
        return {"head_fwd": v[0], "embed_pack_fwd": v[1], "lstm_fwd": sum(v[2:10]), "vocab_ce_fwd": v[10],
                "vocab_ce_bwd": v[11], "lstm_bwd": sum(v[12:20]), "embed_pack_bwd": v[20], "head_bwd": v[21]}

    def dp_adam_shard(self, mc_g, mc_p, lo, hi, step, lr, betas, eps, grad_clip, max_blocks=0):
        """The data-parallel exchange fused with clip_gradient + Adam on this rank's shard flat[lo:hi] (include/snt_b200.h:
        snt_dp_adam_shard): mc_g / mc_p are the multicast addresses of the flat gradient / parameter buffers."""
        f = self.flat
        if hi <= lo:
            return
        _lib.call("snt_dp_adam_shard", C.c_void_p(mc_g), C.c_void_p(mc_p), C.c_void_p(f.p.data_ptr()),
                  C.c_void_p(f.m.data_ptr()), C.c_void_p(f.v.data_ptr()), int(lo), int(hi), float(lr), float(betas[0]),
                  float(betas[1]), float(eps), float(grad_clip if grad_clip is not None else 0.0), 1.0, int(step),
                  int(max_blocks), _lib.stream_ptr())

    def adam(self, lo, hi, step, lr, betas, eps, grad_clip):
        """clip_gradient + Adam (train.py:88-91,145-146) on flat[lo:hi]."""
        f = self.flat
        if hi <= lo:
            return
        P = lambda t: C.c_void_p(t.data_ptr() + 4 * lo)
        _lib.call("snt_clamp_adam", P(f.p), P(f.g), P(f.m), P(f.v), hi - lo, float(lr), float(betas[0]),
                  float(betas[1]), float(eps), float(grad_clip if grad_clip is not None else 0.0), 1.0, int(step),
                  _lib.stream_ptr())
