"""CPU oracle for the Show-and-Tell caption-decoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the arithmetic the reference performs on its hot path
(`/root/reference/models.py:9-67` plus the loss/step glue at `/root/reference/train.py:134-146`).
The arithmetic itself lives in the reference's un-vendored, un-pinned dependency PyTorch
(era torch 0.1.12; here torch 2.11.0): `nn.Embedding`, `nn.LSTM`, `nn.Linear`, `nn.BatchNorm1d`,
`pack_padded_sequence`, `nn.CrossEntropyLoss`, autograd, `optim.Adam`.  The published algorithms of
those layers are restated below; every function cites the reference call site it follows.

PINNING: the reference ships no tests, fixtures or golden vectors for this path (SURVEY.md §8c,
"parity unpinned by the reference").  The oracle is therefore pinned against *outputs of the
reference itself run in the build container*: `oracle/make_golden.py` imports the unmodified
`/root/reference/models.py`, runs it under torch 2.11 CPU and commits the vectors under
`tests/golden/`; `tests/test_oracle_golden.py` checks this restatement against them.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product (`show-and-tell_b200/`) never does and has no CPU fallback.
"""
from __future__ import annotations

import numpy as np

SAMPLE_STEPS = 20  # models.py:60  `for i in range(20)`


# --------------------------------------------------------------------------------------
# packing  (torch.nn.utils.rnn.pack_padded_sequence, called at models.py:51, train.py:135, eval.py:91)
# --------------------------------------------------------------------------------------
def pack_info(lengths):
    """batch_sizes / offsets of a time-major packed sequence.  lengths must be sorted descending
    (data_loader.py:50 sorts the batch; pack_padded_sequence enforces it)."""
    lengths = [int(l) for l in lengths]
    if len(lengths) == 0:
        raise ValueError("empty batch")
    if any(lengths[i] < lengths[i + 1] for i in range(len(lengths) - 1)):
        raise ValueError("lengths must be sorted in decreasing order")
    if lengths[-1] < 1:
        raise ValueError("all lengths must be >= 1")
    T = lengths[0]
    la = np.asarray(lengths)
    bs = np.array([(la > t).sum() for t in range(T)], dtype=np.int64)
    off = np.zeros(T + 1, dtype=np.int64)
    off[1:] = np.cumsum(bs)
    return T, bs, off


def pack_rows(padded, lengths):
    """pack_padded_sequence(padded, lengths, batch_first=True)[0] for padded[B, Tp, ...]."""
    T, bs, off = pack_info(lengths)
    out = np.empty((int(off[-1]),) + padded.shape[2:], dtype=padded.dtype)
    for t in range(T):
        out[off[t]:off[t + 1]] = padded[: bs[t], t]
    return out


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _lstm_weights(params, k):
    return (params[f"lstm.weight_ih_l{k}"], params[f"lstm.weight_hh_l{k}"],
            params[f"lstm.bias_ih_l{k}"], params[f"lstm.bias_hh_l{k}"])


def num_layers(params):
    k = 0
    while f"lstm.weight_ih_l{k}" in params:
        k += 1
    return k


def cast_params(params, dtype):
    return {k: np.asarray(v).astype(dtype) for k, v in params.items()}


# --------------------------------------------------------------------------------------
# encoder head  (models.py:16-17, 27-28: resnet.fc Linear + BatchNorm1d(momentum=0.01))
# --------------------------------------------------------------------------------------
def head_forward(hp, pooled, training=True, momentum=0.01, eps=1e-5):
    """pooled[B,2048] -> features[B,E].  hp: fc.weight[E,2048], fc.bias[E], bn.weight, bn.bias,
    bn.running_mean, bn.running_var.  Returns (features, cache, new_running_mean, new_running_var)."""
    W, b = hp["resnet.fc.weight"], hp["resnet.fc.bias"]
    y = pooled @ W.T + b
    B = y.shape[0]
    if training:
        mu = y.mean(0)
        var = ((y - mu) ** 2).mean(0)  # biased, used for normalisation
        rm = (1 - momentum) * hp["bn.running_mean"] + momentum * mu
        rv = (1 - momentum) * hp["bn.running_var"] + momentum * var * (B / max(B - 1, 1))
    else:
        mu, var = hp["bn.running_mean"], hp["bn.running_var"]
        rm, rv = hp["bn.running_mean"], hp["bn.running_var"]
    rstd = 1.0 / np.sqrt(var + eps)
    yhat = (y - mu) * rstd
    out = yhat * hp["bn.weight"] + hp["bn.bias"]
    cache = dict(pooled=pooled, yhat=yhat, rstd=rstd, training=training)
    return out, cache, rm, rv


def head_backward(hp, cache, dout):
    """Gradients of the head w.r.t. its trainable parameters (the backbone is frozen, models.py:14-15)."""
    yhat, rstd = cache["yhat"], cache["rstd"]
    g = {}
    g["bn.weight"] = (dout * yhat).sum(0)
    g["bn.bias"] = dout.sum(0)
    dyhat = dout * hp["bn.weight"]
    if cache["training"]:
        dy = rstd * (dyhat - dyhat.mean(0) - yhat * (dyhat * yhat).mean(0))
    else:
        dy = dyhat * rstd
    g["resnet.fc.weight"] = dy.T @ cache["pooled"]
    g["resnet.fc.bias"] = dy.sum(0)
    return g


# --------------------------------------------------------------------------------------
# decoder forward  (models.py:47-54)
# --------------------------------------------------------------------------------------
def decoder_inputs(params, features, captions, lengths):
    """models.py:49-51: embed -> cat(feature as step 0) -> pack.  Returns X[N,E] time-major packed."""
    T, bs, off = pack_info(lengths)
    if T > captions.shape[1] + 1:
        raise ValueError("max(lengths) exceeds captions.shape[1] + 1")
    W_emb = params["embed.weight"]
    X = np.empty((int(off[-1]), W_emb.shape[1]), dtype=W_emb.dtype)
    for t in range(T):
        if t == 0:
            X[off[0]:off[1]] = features[: bs[0]]
        else:
            X[off[t]:off[t + 1]] = W_emb[captions[: bs[t], t - 1]]
    return X


def lstm_forward(params, X, lengths):
    """models.py:52 `self.lstm(packed)`; gate equations torch/nn/modules/rnn.py (i,f,g,o row blocks),
    h0 = c0 = 0.  Returns (Hs_last_layer[N,H], caches per layer)."""
    T, bs, off = pack_info(lengths)
    L = num_layers(params)
    caches = []
    inp = X
    for k in range(L):
        W_ih, W_hh, b_ih, b_hh = _lstm_weights(params, k)
        H = W_hh.shape[1]
        N = inp.shape[0]
        Gx = inp @ W_ih.T + (b_ih + b_hh)
        Hs = np.zeros((N, H), dtype=inp.dtype)
        Cs = np.zeros((N, H), dtype=inp.dtype)
        Hprev = np.zeros((N, H), dtype=inp.dtype)
        Cprev = np.zeros((N, H), dtype=inp.dtype)
        act = np.zeros((N, 4 * H), dtype=inp.dtype)
        h = np.zeros((bs[0], H), dtype=inp.dtype)
        c = np.zeros((bs[0], H), dtype=inp.dtype)
        for t in range(T):
            b = int(bs[t])
            h, c = h[:b], c[:b]
            g = Gx[off[t]:off[t + 1]] + h @ W_hh.T
            i_, f_, g_, o_ = (sigmoid(g[:, :H]), sigmoid(g[:, H:2 * H]),
                              np.tanh(g[:, 2 * H:3 * H]), sigmoid(g[:, 3 * H:]))
            Hprev[off[t]:off[t + 1]] = h
            Cprev[off[t]:off[t + 1]] = c
            c = f_ * c + i_ * g_
            h = o_ * np.tanh(c)
            act[off[t]:off[t + 1]] = np.concatenate([i_, f_, g_, o_], 1)
            Hs[off[t]:off[t + 1]] = h
            Cs[off[t]:off[t + 1]] = c
        caches.append(dict(inp=inp, act=act, Cs=Cs, Hprev=Hprev, Cprev=Cprev, Hs=Hs))
        inp = Hs
    return inp, caches


def decoder_forward(params, features, captions, lengths):
    """DecoderRNN.forward (models.py:47-54): returns logits[N,V] and the cache for backward."""
    X = decoder_inputs(params, features, captions, lengths)
    Hs, caches = lstm_forward(params, X, lengths)
    logits = Hs @ params["linear.weight"].T + params["linear.bias"]
    return logits, dict(X=X, Hs=Hs, lstm=caches)


# --------------------------------------------------------------------------------------
# loss  (train.py:53,143 / eval.py:95: nn.CrossEntropyLoss(), reduction = mean over the N packed rows)
# --------------------------------------------------------------------------------------
def cross_entropy(logits, targets):
    m = logits.max(1, keepdims=True)
    lse = (m + np.log(np.exp(logits - m).sum(1, keepdims=True)))[:, 0]
    n = np.arange(logits.shape[0])
    loss = (lse - logits[n, targets]).mean()
    return loss, lse


def cross_entropy_backward(logits, lse, targets, grad_out=1.0):
    N = logits.shape[0]
    d = np.exp(logits - lse[:, None])
    d[np.arange(N), targets] -= 1.0
    return d * (grad_out / N)


# --------------------------------------------------------------------------------------
# decoder backward  (train.py:144 `loss.backward()` through models.py:49-53)
# --------------------------------------------------------------------------------------
def decoder_backward(params, cache, captions, lengths, dlogits, n_features):
    """Manual BPTT.  Returns (grads dict keyed like state_dict, dfeatures[B,E])."""
    T, bs, off = pack_info(lengths)
    L = num_layers(params)
    g = {}
    Hs = cache["Hs"]
    g["linear.weight"] = dlogits.T @ Hs
    g["linear.bias"] = dlogits.sum(0)
    dOut = dlogits @ params["linear.weight"]  # [N,H] gradient into the last layer's hiddens
    for k in reversed(range(L)):
        W_ih, W_hh, _, _ = _lstm_weights(params, k)
        H = W_hh.shape[1]
        lc = cache["lstm"][k]
        act, Cs, Cprev, Hprev = lc["act"], lc["Cs"], lc["Cprev"], lc["Hprev"]
        dG = np.zeros_like(act)
        dh_next = np.zeros((bs[0], H), dtype=act.dtype)
        dc_next = np.zeros((bs[0], H), dtype=act.dtype)
        for t in reversed(range(T)):
            b = int(bs[t])
            sl = slice(off[t], off[t + 1])
            i_, f_, g_, o_ = act[sl, :H], act[sl, H:2 * H], act[sl, 2 * H:3 * H], act[sl, 3 * H:]
            tc = np.tanh(Cs[sl])
            dh = dOut[sl] + dh_next[:b]
            dc = dc_next[:b] + dh * o_ * (1 - tc * tc)
            di = dc * g_ * i_ * (1 - i_)
            df = dc * Cprev[sl] * f_ * (1 - f_)
            dg = dc * i_ * (1 - g_ * g_)
            do = dh * tc * o_ * (1 - o_)
            dGt = np.concatenate([di, df, dg, do], 1)
            dG[sl] = dGt
            dh_next = np.zeros((bs[0], H), dtype=act.dtype)
            dc_next = np.zeros((bs[0], H), dtype=act.dtype)
            dh_next[:b] = dGt @ W_hh
            dc_next[:b] = dc * f_
        g[f"lstm.weight_ih_l{k}"] = dG.T @ lc["inp"]
        g[f"lstm.weight_hh_l{k}"] = dG.T @ Hprev
        g[f"lstm.bias_ih_l{k}"] = dG.sum(0)
        g[f"lstm.bias_hh_l{k}"] = dG.sum(0)
        dOut = dG @ W_ih  # gradient into this layer's input rows
    dX = dOut
    W_emb = params["embed.weight"]
    dW_emb = np.zeros_like(W_emb)
    dfeat = np.zeros((n_features, W_emb.shape[1]), dtype=W_emb.dtype)
    dfeat[: bs[0]] = dX[off[0]:off[1]]
    for t in range(1, T):
        np.add.at(dW_emb, captions[: bs[t], t - 1], dX[off[t]:off[t + 1]])
    g["embed.weight"] = dW_emb
    return g, dfeat


def train_step(params, features, captions, lengths, targets):
    """fwd + CE + bwd exactly as train.py:139-144 would run DecoderRNN.  Returns dict of results."""
    logits, cache = decoder_forward(params, features, captions, lengths)
    loss, lse = cross_entropy(logits, targets)
    dlogits = cross_entropy_backward(logits, lse, targets)
    grads, dfeat = decoder_backward(params, cache, captions, lengths, dlogits, features.shape[0])
    return dict(loss=loss, lse=lse, logits=logits, grads=grads, dfeatures=dfeat, cache=cache)


# --------------------------------------------------------------------------------------
# greedy decode  (models.py:56-67, with the pre-0.2 keepdim semantics of `max(1)[1]`)
# --------------------------------------------------------------------------------------
def greedy_sample(params, features, states=None, steps=SAMPLE_STEPS, forced_ids=None,
                  return_margins=False):
    """ids[B,steps] int64; first-index argmax (torch.max tie-break).  `forced_ids[B,steps]`, when given,
    teacher-forces the token fed to the next step (used by the margin-gated parity check)."""
    L = num_layers(params)
    B = features.shape[0]
    dt = features.dtype
    H = params["lstm.weight_hh_l0"].shape[1]
    if states is None:
        h = [np.zeros((B, H), dt) for _ in range(L)]
        c = [np.zeros((B, H), dt) for _ in range(L)]
    else:
        h = [np.array(states[0][k], dtype=dt) for k in range(L)]
        c = [np.array(states[1][k], dtype=dt) for k in range(L)]
    x = features
    ids = np.zeros((B, steps), dtype=np.int64)
    margins = np.zeros((B, steps), dtype=np.float64)
    for s in range(steps):
        inp = x
        for k in range(L):
            W_ih, W_hh, b_ih, b_hh = _lstm_weights(params, k)
            g = inp @ W_ih.T + b_ih + h[k] @ W_hh.T + b_hh
            i_, f_, g_, o_ = (sigmoid(g[:, :H]), sigmoid(g[:, H:2 * H]),
                              np.tanh(g[:, 2 * H:3 * H]), sigmoid(g[:, 3 * H:]))
            c[k] = f_ * c[k] + i_ * g_
            h[k] = o_ * np.tanh(c[k])
            inp = h[k]
        logits = inp @ params["linear.weight"].T + params["linear.bias"]
        pred = logits.argmax(1)  # numpy argmax returns the first maximal index, like torch.max
        ids[:, s] = pred
        if return_margins:
            part = np.partition(logits, -2, axis=1)
            margins[:, s] = part[:, -1] - part[:, -2]
        nxt = pred if forced_ids is None else forced_ids[:, s]
        x = params["embed.weight"][nxt]
    if return_margins:
        return ids, margins
    return ids


# --------------------------------------------------------------------------------------
# caller-side tail of sample()  (eval.py:101-109)
# --------------------------------------------------------------------------------------
def trim_captions(ids, end_id=2, pad_id=0):
    """eval.py:103-109: for every sampled row, the words kept are those before the first `<end>`
    (`if word == '<end>': break`); a row without `<end>` keeps all its words.  -> (ids with the dropped positions
    replaced by pad_id, lengths[B] int32 = number of words kept).  <end> = 2, <pad> = 0: preprocess.py:75-78."""
    ids = np.asarray(ids, dtype=np.int64)
    out = np.full_like(ids, pad_id)
    lengths = np.zeros(ids.shape[0], dtype=np.int32)
    for b, sentence_ids in enumerate(ids):
        kept = []
        for word_id in sentence_ids:
            if word_id == end_id:
                break
            kept.append(word_id)
        lengths[b] = len(kept)
        out[b, :len(kept)] = kept
    return out, lengths


# --------------------------------------------------------------------------------------
# clip + Adam  (train.py:88-91 clip_gradient clamps each grad to +-grad_clip; train.py:56 optim.Adam)
# --------------------------------------------------------------------------------------
def clamp_adam(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, grad_clip=0.1):
    """One Adam step (torch.optim.Adam defaults, no weight decay, no amsgrad) on clamped grads.
    `step` is the 1-based step count after this update.  Returns new (p, m, v)."""
    g = np.clip(g, -grad_clip, grad_clip) if grad_clip is not None else g
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v
