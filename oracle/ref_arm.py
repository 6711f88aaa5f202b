"""The reference's own `models.py`, timed on host cores.  TEST / BENCH INFRASTRUCTURE ONLY.

`build_ref()` is the committed recipe: in the build container (where `/root/reference` exists) it places an UNMODIFIED,
byte-identical copy of `/root/reference/models.py` under `oracle/_ref/` (git-ignored, so it never enters history, but
not gpurun-ignored, so it travels to the GPU box like the built `.so`).  Nothing under `oracle/_ref/` is edited; a
sha256 of the file is recorded next to it and re-checked on load.

`load()` imports that file the way `oracle/make_golden.py` imports the reference: the only shim is
`torchvision.models.resnet152(pretrained=True)` -> `weights=None` (`models.py:13` wants to download the trunk, and there
is no network; the trunk is never run on this path — the step starts from precomputed pooled features).

`time_full_train` drives the reference's `EncoderCNN` head (`resnet.fc` + `bn`, `models.py:16-17, 27-28`) and its
`DecoderRNN.forward` (`models.py:47-54`) through the step of `train.py:134-146`; `time_greedy` runs the loop of
`DecoderRNN.sample` (`models.py:56-67`) on the reference's own layers with `max(1, keepdim=True)` — the 2017 semantics
the file was written for; under torch >= 0.2 `sample()` as written fails on its second iteration.

The product never imports this (tests/test_cabi_cpu.py::test_product_never_imports_oracle).
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys
import time

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/models.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_FILE = os.path.join(REF_DIR, "models.py")
REF_SHA = os.path.join(REF_DIR, "models.py.sha256")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def build_ref() -> str | None:
    """Copy the reference file where it can travel.  Returns the path, or None when the reference tree is absent."""
    if not os.path.isfile(REF_SRC):
        return REF_FILE if available() else None
    os.makedirs(REF_DIR, exist_ok=True)
    shutil.copyfile(REF_SRC, REF_FILE)
    with open(REF_SHA, "w") as f:
        f.write(f"{_sha(REF_FILE)}  {REF_SRC}\n")
    return REF_FILE


def available() -> bool:
    return os.path.isfile(REF_FILE) and os.path.isfile(REF_SHA)


_MOD = None


def load():
    """Import oracle/_ref/models.py (unmodified) as a module."""
    global _MOD
    if _MOD is not None:
        return _MOD
    if not available():
        raise FileNotFoundError("oracle/_ref/models.py is absent: run oracle.ref_arm.build_ref() in the build container")
    want = open(REF_SHA).read().split()[0]
    if _sha(REF_FILE) != want:
        raise RuntimeError("oracle/_ref/models.py differs from the recorded reference file")
    spec = importlib.util.spec_from_file_location("snt_reference_models", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["snt_reference_models"] = mod
    spec.loader.exec_module(mod)
    _MOD = mod
    return mod


def what() -> str:
    return (f"torch {torch.__version__} CPU, the reference's unmodified models.py (oracle/_ref, sha256 "
            f"{open(REF_SHA).read().split()[0][:12]})")


def make_models(E, H, V, L, seed=0):
    """(EncoderCNN, DecoderRNN) of the reference, random-init; the ResNet trunk is built but never run."""
    ref = load()
    import torchvision.models as tvm
    orig = tvm.resnet152
    tvm.resnet152 = lambda pretrained=False, **kw: orig(weights=None)   # models.py:13, no network here
    try:
        torch.manual_seed(seed)
        enc = ref.EncoderCNN(E)
        dec = ref.DecoderRNN(E, H, V, L)
    finally:
        tvm.resnet152 = orig
    return enc, dec


def head(enc, pooled):
    """EncoderCNN.forward after the frozen trunk (models.py:25-29): the replaced fc, then BatchNorm."""
    return enc.bn(enc.resnet.fc(pooled))


def full_step(enc, dec, opt, criterion, pooled, captions, lengths, targets, grad_clip=0.1):
    """train.py:137-146 on precomputed pooled features."""
    enc.zero_grad()
    dec.zero_grad()
    outputs = dec(head(enc, pooled), captions, lengths)
    loss = criterion(outputs, targets)
    loss.backward()
    for group in opt.param_groups:              # clip_gradient, train.py:88-91
        for p in group["params"]:
            if p.grad is not None:
                p.grad.data.clamp_(-grad_clip, grad_clip)
    opt.step()
    return loss


def time_full_train(B, E, H, V, L, batch, steps=3, warmup=1, threads=None):
    """-> (captions/s, seconds per step, loss); same contract as oracle.torch_port.time_full_train."""
    if threads:
        torch.set_num_threads(threads)
    enc, dec = make_models(E, H, V, L)
    if batch["pooled"].shape[1] != enc.resnet.fc.in_features:
        raise ValueError("the reference's head takes ResNet-152's 2048 pooled features")
    params = [p for p in list(enc.resnet.fc.parameters()) + list(enc.bn.parameters()) + list(dec.parameters())
              if p.requires_grad]                                      # train.py:55-56 (the trunk is frozen, models.py:14-15)
    opt = torch.optim.Adam(params, lr=1e-3)
    criterion = nn.CrossEntropyLoss()                                  # train.py:53
    p = torch.from_numpy(batch["pooled"])
    c = torch.from_numpy(batch["captions"])
    t = torch.from_numpy(batch["targets"])
    ts, loss = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = full_step(enc, dec, opt, criterion, p, c, batch["lengths"], t)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return B / dt, dt, float(loss.detach())


def time_configs0(E=512, H=1024, V=10000, L=1, B=128, steps=1, warmup=1, threads=None, seed=0):
    """BASELINE configs[0]: the reference's whole model on the CPU - frozen ResNet-152 trunk -> Linear + BN head ->
    1-layer LSTM decoder at config.py's sizes (config.py:17,27-29), batch 128 synthetic 224x224 images, captions <= 20
    tokens - one teacher-forced train step as train.py:134-146 runs it (EncoderCNN.forward, models.py:25-29, unmodified).
    -> (images/s, seconds per step, loss)."""
    if threads:
        torch.set_num_threads(threads)
    enc, dec = make_models(E, H, V, L, seed=seed)
    enc.train()
    dec.train()
    rng = np.random.default_rng(seed + 1)
    images = torch.from_numpy(rng.standard_normal((B, 3, 224, 224)).astype(np.float32))
    lengths = np.sort(rng.integers(6, 21, size=B))[::-1].copy()
    lengths[0] = 20
    width = int(lengths[0]) - 1                                  # captions[:, :-1] of train.py:139
    captions = torch.from_numpy(rng.integers(0, V, size=(B, width)).astype(np.int64))
    lengths = [int(x) for x in lengths]
    from torch.nn.utils.rnn import pack_padded_sequence
    tgt_src = torch.from_numpy(rng.integers(0, V, size=(B, width + 1)).astype(np.int64))
    targets = pack_padded_sequence(tgt_src, lengths, batch_first=True)[0]       # train.py:135
    params = [p for p in list(enc.parameters()) + list(dec.parameters()) if p.requires_grad]   # train.py:55-56
    opt = torch.optim.Adam(params, lr=1e-3)
    criterion = nn.CrossEntropyLoss()
    ts, loss = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        enc.zero_grad()
        dec.zero_grad()
        outputs = dec(enc(images), captions, lengths)            # train.py:139 through models.py:25-29 and :47-54
        loss = criterion(outputs, targets)
        loss.backward()
        for group in opt.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    p.grad.data.clamp_(-0.1, 0.1)
        opt.step()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return B / dt, dt, float(loss.detach())


@torch.no_grad()
def sample_keepdim(dec, features, states=None, steps=20):
    """models.py:56-67 on the reference's layers, `max(1, keepdim=True)` restoring the 2017 shape of `predicted`."""
    ids = []
    inputs = features.unsqueeze(1)
    for _ in range(steps):
        hiddens, states = dec.lstm(inputs, states)
        outputs = dec.linear(hiddens.squeeze(1))
        predicted = outputs.max(1, keepdim=True)[1]
        ids.append(predicted)
        inputs = dec.embed(predicted)
    return torch.cat(ids, 1)


def time_greedy(B, E, H, V, L, features, steps=2, warmup=1, threads=None):
    """-> (tokens/s, seconds per sample() of B features x 20 tokens)."""
    if threads:
        torch.set_num_threads(threads)
    ref = load()
    torch.manual_seed(0)
    dec = ref.DecoderRNN(E, H, V, L).eval()
    f = torch.from_numpy(features)
    for _ in range(warmup):
        sample_keepdim(dec, f)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        sample_keepdim(dec, f)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return B * 20 / dt, dt


if __name__ == "__main__":
    print(build_ref())
