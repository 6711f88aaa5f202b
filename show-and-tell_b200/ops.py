"""Autograd-aware host side of the caption-decoder path: each function here replaces the PyTorch calls the
reference makes in `models.py` / `train.py` with calls into the C-ABI library (`_lib.py`).

  head(...)            <- EncoderCNN.forward's  self.bn(self.resnet.fc(pooled))          models.py:27-28
  decoder_logits(...)  <- DecoderRNN.forward                                             models.py:47-54
  decoder_loss(...)    <- DecoderRNN.forward + nn.CrossEntropyLoss()(outputs, targets)   train.py:139-143
  greedy(...)          <- DecoderRNN.sample                                              models.py:56-67
  clamp_adam_(...)     <- clip_gradient + optim.Adam.step                                train.py:88-91,145-146

All tensors must live on a CUDA device; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import PREC, call, host_i32, ptr, require_cuda, stream_ptr, workspace

SAMPLE_STEPS = 20  # models.py:60

# SNT_CE_RECOMPUTE=1: the memory-lean loss backward that recomputes the logits per chunk instead of keeping the bf16 softmax
# numerators of the whole batch (N x V x 2 bytes) from the forward pass.
CE_RECOMPUTE = os.environ.get("SNT_CE_RECOMPUTE", "0") == "1"

def _act_dtype(prec):
    return torch.float32 if prec == "fp32" else torch.bfloat16


def batch_sizes_from_lengths(lengths, max_steps=None):
    """`lengths` as the reference passes them (list of ints sorted descending, data_loader.py:50) ->
    int32 batch_sizes[T] of the packed sequence (what pack_padded_sequence computes at models.py:51)."""
    key = (tuple(lengths) if isinstance(lengths, (list, tuple)) else None, max_steps)
    if key[0] is not None:
        hit = _BS_CACHE.get(key)
        if hit is not None:
            return hit
    bs = _batch_sizes_from_lengths(lengths, max_steps)
    if key[0] is not None:
        if len(_BS_CACHE) > 64:
            _BS_CACHE.clear()
        bs.setflags(write=False)
        _BS_CACHE[key] = bs
    return bs


_BS_CACHE = {}   # lengths tuple -> batch_sizes (the same validation + numpy pass costs ~140 us per call otherwise)


def _batch_sizes_from_lengths(lengths, max_steps=None):
    l = np.asarray([int(x) for x in lengths], dtype=np.int64)
    if l.ndim != 1 or l.size == 0:
        raise RuntimeError("lengths must be a non-empty 1-D sequence")
    if (l[:-1] < l[1:]).any():
        raise RuntimeError("`lengths` array must be sorted in decreasing order")  # torch's message, enforce_sorted
    if l[-1] < 1:
        raise RuntimeError("Length of all samples has to be greater than 0, but found an element in "
                           "'lengths' that is <= 0")
    T = int(l[0])
    if max_steps is not None and T > max_steps:
        raise RuntimeError(f"max(lengths)={T} exceeds the {max_steps} timesteps available (captions.shape[1] + 1)")
    bs = (l[None, :] > np.arange(T)[:, None]).sum(1).astype(np.int32)
    return bs


# ---------------------------------------------------------------------------------------------------------
# encoder head
# ---------------------------------------------------------------------------------------------------------
class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pooled, w_fc, b_fc, gamma, beta, running_mean, running_var, training, momentum, eps, prec):
        require_cuda(pooled, w_fc, b_fc, gamma, beta, running_mean, running_var)
        pooled = pooled.contiguous().float()
        B, K = pooled.shape
        E = w_fc.shape[0]
        dev = pooled.device
        feats = torch.empty(B, E, device=dev)
        yhat = torch.empty(B, E, device=dev)
        rstd = torch.empty(E, device=dev)
        p = PREC[prec]
        nb = _lib.lib().snt_head_workspace_bytes(p, B, K, E)
        ws = workspace(nb, dev)
        call("snt_head_fwd", p, ptr(pooled), ptr(w_fc), ptr(b_fc), ptr(gamma), ptr(beta), ptr(running_mean),
             ptr(running_var), 1 if training else 0, float(momentum), float(eps), B, K, E, ptr(feats), ptr(yhat),
             ptr(rstd), ptr(ws), ws.numel(), stream_ptr())
        ctx.save_for_backward(pooled, yhat, rstd, gamma)
        ctx.meta = (training, prec)
        return feats

    @staticmethod
    def backward(ctx, dfeat):
        pooled, yhat, rstd, gamma = ctx.saved_tensors
        training, prec = ctx.meta
        B, K = pooled.shape
        E = yhat.shape[1]
        dev = pooled.device
        dfeat = dfeat.contiguous().float()
        d_w = torch.empty(E, K, device=dev)
        d_b = torch.empty(E, device=dev)
        d_g = torch.empty(E, device=dev)
        d_be = torch.empty(E, device=dev)
        p = PREC[prec]
        nb = _lib.lib().snt_head_workspace_bytes(p, B, K, E)
        ws = workspace(nb, dev)
        call("snt_head_bwd", p, ptr(dfeat), ptr(pooled), ptr(yhat), ptr(rstd), ptr(gamma), 1 if training else 0,
             B, K, E, ptr(d_w), ptr(d_b), ptr(d_g), ptr(d_be), ptr(ws), ws.numel(), stream_ptr())
        return None, d_w, d_b, d_g, d_be, None, None, None, None, None, None


def head(pooled, w_fc, b_fc, gamma, beta, running_mean, running_var, training, momentum=0.01, eps=1e-5,
         prec="fp32"):
    """features[B,E] = BatchNorm1d(Linear(pooled[B,2048]))  (models.py:27-28); updates the running stats in
    place when `training`."""
    return _Head.apply(pooled, w_fc, b_fc, gamma, beta, running_mean, running_var, bool(training), momentum, eps, prec)


# ---------------------------------------------------------------------------------------------------------
# decoder: gather/concat/pack + L LSTM layers (shared by the logits and the fused-loss functions)
# ---------------------------------------------------------------------------------------------------------
class _Hidden:
    """Tensors saved between forward and backward of the recurrent part."""
    __slots__ = ("prec", "bs", "bs_ptr", "T", "N", "B", "E", "H", "V", "L", "captions", "layers", "x")


def _hidden_fwd(prec, features, captions, bs, w_emb, lstm_w):
    """-> (hs_last (act) [N,H], saved).  lstm_w: list of (w_ih, w_hh, b_ih, b_hh) per layer."""
    require_cuda(features, captions, w_emb)
    dev = features.device
    p = PREC[prec]
    act = _act_dtype(prec)
    B, E = features.shape
    V = w_emb.shape[0]
    T = int(bs.shape[0])
    N = int(bs.sum())
    if int(bs[0]) != B:
        raise RuntimeError(f"len(lengths)={int(bs[0])} does not match features.shape[0]={B}")
    bs_arr, bs_ptr = host_i32(bs)
    x = torch.empty(N, E, device=dev, dtype=act)
    call("snt_embed_pack_fwd", ptr(features), ptr(w_emb), ptr(captions), captions.stride(0) if captions.numel() else 0,
         bs_ptr, T, E, V, ptr(x) if prec == "fp32" else None, ptr(x) if prec == "bf16" else None, stream_ptr())
    s = _Hidden()
    s.prec, s.bs, s.bs_ptr, s.T, s.N, s.B, s.E, s.V, s.L = prec, bs_arr, bs_ptr, T, N, B, E, V, len(lstm_w)
    s.captions, s.x, s.layers = captions, x, []
    inp, in_dim = x, E
    H = lstm_w[0][1].shape[1]
    s.H = H
    for (w_ih, w_hh, b_ih, b_hh) in lstm_w:
        gates = torch.empty(N, 4 * H, device=dev)
        cs = torch.empty(N, H, device=dev)
        hs = torch.empty(N, H, device=dev, dtype=act)
        hprev = torch.empty(N, H, device=dev, dtype=act)
        nb = _lib.lib().snt_lstm_workspace_bytes(p, N, B, in_dim, H)
        ws = workspace(nb, dev)
        call("snt_lstm_fwd", p, ptr(inp), in_dim, H, ptr(w_ih), ptr(w_hh), ptr(b_ih), ptr(b_hh), bs_ptr, T,
             ptr(gates), ptr(cs), ptr(hs), ptr(hprev), ptr(ws), ws.numel(), stream_ptr())
        s.layers.append((inp, in_dim, gates, cs, hs, hprev))
        inp, in_dim = hs, H
    return inp, s


def _hidden_bwd(s, d_hs, w_emb, lstm_w, need_dfeat):
    """BPTT through the layers (top to bottom), then the gather's backward.  `gates` buffers are consumed.
    -> (dfeatures | None, d_w_emb, [(d_w_ih, d_w_hh, d_b_ih, d_b_hh)] per layer)."""
    dev = d_hs.device
    p = PREC[s.prec]
    H = s.H
    grads = [None] * s.L
    for k in reversed(range(s.L)):
        inp, in_dim, gates, cs, hs, hprev = s.layers[k]
        w_ih, w_hh, _, _ = lstm_w[k]
        d_w_ih = torch.empty_like(w_ih)
        d_w_hh = torch.empty_like(w_hh)
        d_b = torch.empty(4 * H, device=dev)
        dx = torch.empty(s.N, in_dim, device=dev)
        nb = _lib.lib().snt_lstm_workspace_bytes(p, s.N, s.B, in_dim, H)
        ws = workspace(nb, dev)
        call("snt_lstm_bwd", p, ptr(d_hs), ptr(gates), ptr(cs), ptr(hprev), ptr(inp), in_dim, H, ptr(w_ih), ptr(w_hh),
             s.bs_ptr, s.T, ptr(d_w_ih), ptr(d_w_hh), ptr(d_b), ptr(dx), ptr(ws), ws.numel(), stream_ptr())
        grads[k] = (d_w_ih, d_w_hh, d_b, d_b.clone())  # b_ih and b_hh receive the same gradient
        d_hs = dx
    dfeat = torch.empty(s.B, s.E, device=dev) if need_dfeat else None
    d_w_emb = torch.empty_like(w_emb)
    cap = s.captions
    nb = _lib.lib().snt_embed_bwd_workspace_bytes(s.N, s.V)
    ws = workspace(nb, dev)
    call("snt_embed_pack_bwd", ptr(d_hs), ptr(cap), cap.stride(0) if cap.numel() else 0, s.bs_ptr, s.T, s.B, s.E,
         s.V, ptr(dfeat), ptr(d_w_emb), ptr(ws), ws.numel(), stream_ptr())
    return dfeat, d_w_emb, grads


def _flatten_lstm(lstm_w):
    return [t for layer in lstm_w for t in layer]


def _unflatten_lstm(flat):
    return [tuple(flat[i:i + 4]) for i in range(0, len(flat), 4)]


def _prep_inputs(features, captions):
    features = features.contiguous().float()
    if captions.dtype != torch.int64:
        captions = captions.long()
    if captions.dim() != 2:
        raise RuntimeError("captions must be [B, Tc]")
    if captions.numel() and captions.stride(1) != 1:
        captions = captions.contiguous()
    if captions.shape[0] != features.shape[0]:
        raise RuntimeError("captions and features disagree on the batch size")
    return features, captions


class _DecoderLogits(torch.autograd.Function):
    """DecoderRNN.forward (models.py:47-54): packed logits [N,V], materialised (strict drop-in)."""

    @staticmethod
    def forward(ctx, features, captions, bs, prec, w_emb, w_out, b_out, *lstm_flat):
        lstm_w = _unflatten_lstm(lstm_flat)
        hs, s = _hidden_fwd(prec, features, captions, bs, w_emb, lstm_w)
        N, H, V = s.N, s.H, w_out.shape[0]
        dev = features.device
        logits = torch.empty(N, V, device=dev)
        p = PREC[prec]
        nb = _lib.lib().snt_linear_workspace_bytes(p, N, H, V)
        ws = workspace(nb, dev)
        call("snt_linear_fwd", p, ptr(hs), ptr(w_out), ptr(b_out), N, H, V, ptr(logits), ptr(ws), ws.numel(),
             stream_ptr())
        ctx.s = s
        ctx.save_for_backward(w_emb, w_out, *lstm_flat)
        ctx.need_dfeat = features.requires_grad
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        s = ctx.s
        w_emb, w_out, *lstm_flat = ctx.saved_tensors
        lstm_w = _unflatten_lstm(lstm_flat)
        dev = dlogits.device
        dlogits = dlogits.contiguous().float()
        N, H, V = s.N, s.H, w_out.shape[0]
        hs = s.layers[-1][4]
        d_hs = torch.empty(N, H, device=dev)
        d_w_out = torch.empty_like(w_out)
        d_b_out = torch.empty(V, device=dev)
        p = PREC[s.prec]
        nb = _lib.lib().snt_linear_workspace_bytes(p, N, H, V)
        ws = workspace(nb, dev)
        call("snt_linear_bwd", p, ptr(dlogits), ptr(hs), ptr(w_out), N, H, V, ptr(d_hs), ptr(d_w_out), ptr(d_b_out),
             ptr(ws), ws.numel(), stream_ptr())
        dfeat, d_w_emb, lg = _hidden_bwd(s, d_hs, w_emb, lstm_w, ctx.need_dfeat)
        ctx.s = None
        return (dfeat, None, None, None, d_w_emb, d_w_out, d_b_out, *_flatten_lstm(lg))


class _DecoderLoss(torch.autograd.Function):
    """DecoderRNN.forward + CrossEntropyLoss (train.py:139-143) with the vocab projection fused into the
    log-softmax / cross-entropy: the [N,V] logits are never materialised as a whole."""

    @staticmethod
    def forward(ctx, features, captions, targets, bs, prec, grad_scale, want_grad, w_emb, w_out, b_out, *lstm_flat):
        lstm_w = _unflatten_lstm(lstm_flat)
        hs, s = _hidden_fwd(prec, features, captions, bs, w_emb, lstm_w)
        N, H, V = s.N, s.H, w_out.shape[0]
        dev = features.device
        if targets.dtype != torch.int64:
            targets = targets.long()
        targets = targets.contiguous()
        if targets.numel() != N:
            raise RuntimeError(f"Expected input batch_size ({N}) to match target batch_size ({targets.numel()}).")
        lse = torch.empty(N, device=dev)
        loss = torch.empty((), device=dev)
        p = PREC[prec]
        # A backward will follow: run the logits contraction once and keep the softmax numerators (bf16 mode; see
        # snt_vocab_ce_train_fwd).  Otherwise (evaluation, fp32 mode, SNT_CE_RECOMPUTE=1): statistics only.
        ctx.stored = prec == "bf16" and want_grad and any(ctx.needs_input_grad) and not CE_RECOMPUTE
        if ctx.stored:
            u = torch.empty(N, (V + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
            inv_s = torch.empty(N, device=dev)
            hs_scaled = torch.empty(N, H, device=dev, dtype=torch.bfloat16)
            w_bf16 = torch.empty(V, H, device=dev, dtype=torch.bfloat16)
            nb = _lib.lib().snt_vocab_ce_train_workspace_bytes(p, N, H, V)
            ws = workspace(nb, dev)
            call("snt_vocab_ce_train_fwd", p, ptr(hs), ptr(w_out), ptr(b_out), ptr(targets), N, H, V, ptr(lse),
                 ptr(loss), ptr(u), ptr(inv_s), ptr(hs_scaled), ptr(w_bf16), ptr(ws), ws.numel(), stream_ptr())
            ctx.ce_saved = (u, inv_s, hs_scaled, w_bf16)
        else:
            nb = _lib.lib().snt_vocab_ce_workspace_bytes(p, N, H, V)
            ws = workspace(nb, dev)
            call("snt_vocab_ce_fwd", p, ptr(hs), ptr(w_out), ptr(b_out), ptr(targets), N, H, V, ptr(lse), ptr(loss),
                 ptr(ws), ws.numel(), stream_ptr())
        ctx.s = s
        ctx.save_for_backward(w_emb, w_out, b_out, targets, lse, *lstm_flat)
        ctx.need_dfeat = features.requires_grad
        ctx.grad_scale = grad_scale
        if grad_scale != 1.0:  # data-parallel: this rank's share of the global-mean loss
            loss = loss * grad_scale
        return loss

    @staticmethod
    def backward(ctx, dloss):
        s = ctx.s
        w_emb, w_out, b_out, targets, lse, *lstm_flat = ctx.saved_tensors
        lstm_w = _unflatten_lstm(lstm_flat)
        dev = dloss.device
        N, H, V = s.N, s.H, w_out.shape[0]
        hs = s.layers[-1][4]
        dloss = dloss.contiguous().float()
        d_hs = torch.empty(N, H, device=dev)
        d_w_out = torch.empty_like(w_out)
        d_b_out = torch.empty(V, device=dev)
        p = PREC[s.prec]
        if ctx.stored:
            u, inv_s, hs_scaled, w_bf16 = ctx.ce_saved
            ctx.ce_saved = None
            nb = _lib.lib().snt_vocab_ce_train_workspace_bytes(p, N, H, V)
            ws = workspace(nb, dev)
            call("snt_vocab_ce_train_bwd", p, ptr(u), ptr(inv_s), ptr(hs_scaled), ptr(w_bf16), ptr(dloss),
                 float(ctx.grad_scale), N, H, V, ptr(d_hs), ptr(d_w_out), ptr(d_b_out), ptr(ws), ws.numel(), stream_ptr())
            del u, inv_s, hs_scaled, w_bf16
        else:
            nb = _lib.lib().snt_vocab_ce_workspace_bytes(p, N, H, V)
            ws = workspace(nb, dev)
            call("snt_vocab_ce_bwd", p, ptr(hs), ptr(w_out), ptr(b_out), ptr(targets), ptr(lse), ptr(dloss),
                 float(ctx.grad_scale), N, H, V, ptr(d_hs), ptr(d_w_out), ptr(d_b_out), ptr(ws), ws.numel(), stream_ptr())
        dfeat, d_w_emb, lg = _hidden_bwd(s, d_hs, w_emb, lstm_w, ctx.need_dfeat)
        ctx.s = None
        return (dfeat, None, None, None, None, None, None, d_w_emb, d_w_out, d_b_out, *_flatten_lstm(lg))


def decoder_logits(features, captions, lengths, w_emb, lstm_w, w_out, b_out, prec="fp32"):
    features, captions = _prep_inputs(features, captions)
    bs = batch_sizes_from_lengths(lengths, captions.shape[1] + 1)
    return _DecoderLogits.apply(features, captions, bs, prec, w_emb, w_out, b_out, *_flatten_lstm(lstm_w))


def decoder_loss(features, captions, lengths, targets, w_emb, lstm_w, w_out, b_out, prec="bf16", grad_scale=1.0):
    """mean_n CE(logits_n, targets_n) over the N packed rows (train.py:53,143)."""
    features, captions = _prep_inputs(features, captions)
    bs = batch_sizes_from_lengths(lengths, captions.shape[1] + 1)
    return _DecoderLoss.apply(features, captions, targets, bs, prec, float(grad_scale), torch.is_grad_enabled(), w_emb,
                              w_out, b_out, *_flatten_lstm(lstm_w))


# ---------------------------------------------------------------------------------------------------------
# greedy decode
# ---------------------------------------------------------------------------------------------------------
@torch.no_grad()
def greedy(features, w_emb, lstm_w, w_out, b_out, states=None, steps=SAMPLE_STEPS, prec="fp32"):
    """ids[B,steps] int64: `steps` iterations of LSTM step -> Linear -> first-index argmax -> embed
    (models.py:59-66).  states: None or (h0, c0), each [L,B,H]."""
    require_cuda(features, w_emb, w_out, b_out)
    features = features.contiguous().float()
    dev = features.device
    B, E = features.shape
    V, H = w_out.shape
    L = len(lstm_w)
    h0 = c0 = None
    if states is not None:
        h0, c0 = states
        h0 = h0.contiguous().float()
        c0 = c0.contiguous().float()
        if tuple(h0.shape) != (L, B, H) or tuple(c0.shape) != (L, B, H):
            raise RuntimeError(f"Expected hidden size {(L, B, H)}, got {tuple(h0.shape)} / {tuple(c0.shape)}")
    ids = torch.empty(B, steps, dtype=torch.int64, device=dev)
    arr = lambda idx: (C.c_void_p * L)(*[w[idx].data_ptr() for w in lstm_w])
    p = PREC[prec]
    nb = _lib.lib().snt_greedy_workspace_bytes(p, B, E, H, V, L)
    ws = workspace(nb, dev)
    call("snt_greedy_decode", p, ptr(features), ptr(w_emb), L, arr(0), arr(1), arr(2), arr(3), ptr(w_out), ptr(b_out),
         ptr(h0), ptr(c0), B, E, H, V, int(steps), ptr(ids), ptr(ws), ws.numel(), stream_ptr())
    return ids


@torch.no_grad()
def trim_captions(ids, end_id=2, pad_id=0, return_ids=True):
    """The caller-side tail of sample() (eval.py:101-109) on the device: -> (ids_trimmed[B,S] int64 or None,
    lengths[B] int32) where lengths[b] = position of the first `end_id` in ids[b] (S if none) = the number of words
    eval.py keeps, and ids_trimmed has every later position replaced by `pad_id`.  `end_id`/`pad_id` default to the
    reference vocabulary's <end> = 2 / <pad> = 0 (preprocess.py:75-78)."""
    require_cuda(ids)
    if ids.dtype != torch.int64:
        raise RuntimeError("trim_captions needs int64 ids (what sample() returns)")
    squeeze = ids.dim() == 1                       # sample() squeezes a batch of one (models.py:66)
    ids2 = (ids.unsqueeze(0) if squeeze else ids).contiguous()
    if ids2.dim() != 2:
        raise RuntimeError("trim_captions expects ids of shape [B,S] (or [S])")
    B, S = ids2.shape
    lengths = torch.empty(B, dtype=torch.int32, device=ids2.device)
    out = torch.empty_like(ids2) if return_ids else None
    call("snt_caption_trim", ptr(ids2), B, int(S), int(end_id), int(pad_id), ptr(lengths), ptr(out), stream_ptr())
    if squeeze and out is not None:
        out = out.squeeze(0)
    return out, lengths


# ---------------------------------------------------------------------------------------------------------
# clip_gradient + Adam
# ---------------------------------------------------------------------------------------------------------
@torch.no_grad()
def clamp_adam_(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_clip=0.1, grad_scale=1.0):
    """In place: g <- clamp(g*grad_scale, +-grad_clip) (train.py:88-91), then one torch.optim.Adam step
    (train.py:56,146).  p, g, m, v: fp32 CUDA tensors of equal numel; `step` is 1-based."""
    require_cuda(p, g, m, v)
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("clamp_adam_ needs contiguous fp32 tensors")
    call("snt_clamp_adam", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(betas[0]), float(betas[1]),
         float(eps), float(grad_clip if grad_clip is not None else 0.0), float(grad_scale), int(step), stream_ptr())


@torch.no_grad()
def clamp_adam_multi_(params, grads, ms, vs, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_clip=0.1,
                      grad_scale=1.0):
    """clamp_adam_ for a list of tensors in ONE kernel launch (train.py:88-91,145-146 over all param groups)."""
    k = len(params)
    if k == 0:
        return
    for t in (*params, *grads, *ms, *vs):
        require_cuda(t)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("clamp_adam_multi_ needs contiguous fp32 tensors")
    arr = lambda ts: (C.c_void_p * k)(*[t.data_ptr() for t in ts])
    sizes = (C.c_int64 * k)(*[p.numel() for p in params])
    call("snt_clamp_adam_multi", k, arr(params), arr(grads), arr(ms), arr(vs), sizes, float(lr), float(betas[0]),
         float(betas[1]), float(eps), float(grad_clip if grad_clip is not None else 0.0), float(grad_scale), int(step),
         stream_ptr())
