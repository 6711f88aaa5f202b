"""Generate golden vectors by running the UNMODIFIED reference (`/root/reference/models.py`).

Run in the build container only (the reference tree does not exist on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz

What is recorded per case: the reference module's own weights (state_dict), the inputs, and the
outputs of the reference code path under torch (CPU, fp32 and an fp64 copy of the same module):
logits, CE loss, every parameter gradient, dfeatures, greedy ids (+ fp64 top-2 margins).
`DecoderRNN.sample` needs one shim: `outputs.max(1)[1]` (models.py:63) relied on the pre-0.2 torch
behaviour of keeping the reduced dim; we subclass and restate that loop with `keepdim=True`
 — the reference file itself is never edited.  The fc+bn head uses the reference
`EncoderCNN` object with torchvision's resnet152 constructed without pretrained weights (no network).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; goldens can only be regenerated in the build container")
    sys.path.insert(0, REF)
    import torchvision.models as tvm
    _orig = tvm.resnet152
    tvm.resnet152 = lambda pretrained=False, **kw: _orig(weights=None)  # models.py:13 wants the network
    import models as ref_models  # noqa
    sys.path.pop(0)
    return ref_models


def sample_keepdim(dec, features, states=None):
    """models.py:56-67 with max(1, keepdim=True) restoring 2017 semantics; also returns top-2 margins."""
    ids, margins = [], []
    inputs = features.unsqueeze(1)
    for _ in range(20):
        hiddens, states = dec.lstm(inputs, states)
        outputs = dec.linear(hiddens.squeeze(1))
        predicted = outputs.max(1, keepdim=True)[1]
        top2 = outputs.topk(2, dim=1)[0]
        margins.append((top2[:, 0] - top2[:, 1]).unsqueeze(1))
        ids.append(predicted)
        inputs = dec.embed(predicted)
    return torch.cat(ids, 1), torch.cat(margins, 1)


def run_decoder(dec, feats, caps, lengths, targets):
    dec.zero_grad()
    feats = feats.clone().requires_grad_(True)
    logits = dec(feats, caps, lengths)                    # models.py:47-54
    loss = nn.CrossEntropyLoss()(logits, targets)         # train.py:53,143
    loss.backward()                                       # train.py:144
    grads = {k: p.grad.detach().numpy().copy() for k, p in dec.named_parameters()}
    return logits.detach().numpy().copy(), float(loss), grads, feats.grad.numpy().copy()


def decoder_case(ref, name, B, E, H, V, L, lengths, convention, seed):
    torch.manual_seed(seed)
    dec = ref.DecoderRNN(E, H, V, L)
    rng = np.random.default_rng(seed + 100)
    lengths = list(lengths)
    tmax = max(lengths)
    caps = np.zeros((B, tmax), dtype=np.int64)
    for i, l in enumerate(lengths):
        caps[i, :l] = rng.integers(4, V, size=l)
        caps[i, 0] = 1
        caps[i, l - 1] = 2
    feats = rng.standard_normal((B, E)).astype(np.float32)
    caps_t = torch.from_numpy(caps)
    if convention == "a":      # eval.py:91-93
        in_caps, in_len = caps_t, lengths
        targets = pack_padded_sequence(caps_t, lengths, batch_first=True)[0]
    else:                      # train.py:134-139
        in_len = [l - 1 for l in lengths]
        in_caps = caps_t[:, :-1]
        targets = pack_padded_sequence(caps_t[:, 1:], in_len, batch_first=True)[0]
    out = dict(B=B, E=E, H=H, V=V, L=L, convention=convention,
               captions=in_caps.numpy().copy(), lengths=np.asarray(in_len), targets=targets.numpy().copy(),
               features=feats)
    for k, v in dec.state_dict().items():
        out["param." + k] = v.numpy().copy()
    for tag, mod, f in (("f32", dec, torch.from_numpy(feats)),
                        ("f64", copy.deepcopy(dec).double(), torch.from_numpy(feats).double())):
        logits, loss, grads, dfeat = run_decoder(mod, f, in_caps, in_len, targets)
        out[f"{tag}.logits"] = logits
        out[f"{tag}.loss"] = np.asarray(loss)
        out[f"{tag}.dfeatures"] = dfeat
        for k, g in grads.items():
            out[f"{tag}.grad.{k}"] = g
        with torch.no_grad():
            ids, margins = sample_keepdim(mod, f)
        out[f"{tag}.greedy_ids"] = ids.numpy().copy()
        out[f"{tag}.greedy_margins"] = margins.numpy().copy()
    # verbatim sample() must fail exactly the way SURVEY.md §0.4 records, which is why the shim exists
    try:
        dec.sample(torch.from_numpy(feats), None)
        out["verbatim_sample_runs"] = np.asarray(1)
    except Exception:
        out["verbatim_sample_runs"] = np.asarray(0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N =", int(sum(in_len)), "loss", float(out["f32.loss"]))


def head_case(ref, name, B, E, seed):
    torch.manual_seed(seed)
    enc = ref.EncoderCNN(E)                                # models.py:10-18 (resnet152 weights=None)
    fc, bn = enc.resnet.fc, enc.bn
    rng = np.random.default_rng(seed + 7)
    pooled = (0.5 * np.abs(rng.standard_normal((B, 2048)))).astype(np.float32)
    bn.weight.data = torch.from_numpy(rng.uniform(0.5, 1.5, E).astype(np.float32))
    bn.bias.data = torch.from_numpy(rng.uniform(-0.5, 0.5, E).astype(np.float32))
    dout = rng.standard_normal((B, E)).astype(np.float32)
    out = dict(B=B, E=E, pooled=pooled, dout=dout)
    for k in ("weight", "bias"):
        out["param.resnet.fc." + k] = getattr(fc, k).detach().numpy().copy()
        out["param.bn." + k] = getattr(bn, k).detach().numpy().copy()
    out["param.bn.running_mean"] = bn.running_mean.numpy().copy()
    out["param.bn.running_var"] = bn.running_var.numpy().copy()
    for mode in ("train", "eval"):
        enc.train(mode == "train")
        for p in (fc.weight, fc.bias, bn.weight, bn.bias):
            p.grad = None
        y = bn(fc(torch.from_numpy(pooled)))               # models.py:27-28 on the pooled features
        y.backward(torch.from_numpy(dout))
        out[f"{mode}.features"] = y.detach().numpy().copy()
        out[f"{mode}.grad.resnet.fc.weight"] = fc.weight.grad.numpy().copy()
        out[f"{mode}.grad.resnet.fc.bias"] = fc.bias.grad.numpy().copy()
        out[f"{mode}.grad.bn.weight"] = bn.weight.grad.numpy().copy()
        out[f"{mode}.grad.bn.bias"] = bn.bias.grad.numpy().copy()
        out[f"{mode}.running_mean_after"] = bn.running_mean.numpy().copy()
        out[f"{mode}.running_var_after"] = bn.running_var.numpy().copy()
        if mode == "train":   # eval case starts from the updated stats
            out["eval.running_mean_before"] = bn.running_mean.numpy().copy()
            out["eval.running_var_before"] = bn.running_var.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok")


def adam_case(name, n, steps, seed):
    """train.py:88-91 (`param.grad.data.clamp_(-grad_clip, grad_clip)`) + train.py:56 optim.Adam(lr)."""
    rng = np.random.default_rng(seed)
    p0 = rng.standard_normal(n).astype(np.float32)
    p = nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p], lr=1e-3)
    gs, ps = [], []
    for _ in range(steps):
        g = (rng.standard_normal(n) * 0.2).astype(np.float32)
        p.grad = torch.from_numpy(g.copy())
        for group in opt.param_groups:
            for param in group["params"]:
                param.grad.data.clamp_(-0.1, 0.1)
        opt.step()
        gs.append(g)
        ps.append(p.detach().numpy().copy())
    st = opt.state[p]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), p0=p0, grads=np.stack(gs), params=np.stack(ps),
                        exp_avg=st["exp_avg"].numpy().copy(), exp_avg_sq=st["exp_avg_sq"].numpy().copy())
    print(name, "ok")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref = import_reference()
    decoder_case(ref, "dec_l1_a", B=6, E=8, H=16, V=23, L=1, lengths=[7, 5, 5, 3, 2, 1], convention="a", seed=0)
    decoder_case(ref, "dec_l2_b", B=5, E=8, H=12, V=19, L=2, lengths=[7, 7, 5, 3, 3], convention="b", seed=1)
    decoder_case(ref, "dec_b1", B=1, E=8, H=16, V=23, L=1, lengths=[4], convention="a", seed=2)
    decoder_case(ref, "dec_l1_mid", B=37, E=24, H=40, V=301, L=1,
                 lengths=sorted(np.random.default_rng(5).integers(2, 15, 37).tolist(), reverse=True),
                 convention="a", seed=3)
    head_case(ref, "head", B=7, E=8, seed=4)
    adam_case("adam", n=64, steps=5, seed=6)


if __name__ == "__main__":
    main()
