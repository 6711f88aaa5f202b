"""CPU baseline of the caption-decoder path on the reference's own substrate.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference's hot path (`/root/reference/models.py:31-67`, `train.py:134-146`) is a composition of torch.nn
layers; `/root/reference` does not exist on the GPU box, so this file restates that composition with the same
torch.nn building blocks (Embedding -> feature as step 0 -> pack_padded_sequence -> LSTM -> Linear ->
CrossEntropyLoss -> backward; and the 20-step greedy loop with the pre-0.2 keepdim semantics of `max(1)[1]`),
to be timed on the GPU box's host cores by bench.py (`cpu_baseline`, `--impl reference`).
PINNED: tests/test_oracle_golden.py::test_torch_port_matches_reference checks it against the vectors the
unmodified reference produced (tests/golden, oracle/make_golden.py).  The product never imports this.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence


class CaptionDecoderCPU(nn.Module):
    def __init__(self, embed, hidden, vocab, layers):
        super().__init__()
        # attribute names follow the reference so that its state_dict loads (models.py:35-37)
        self.embed = nn.Embedding(vocab, embed)
        self.lstm = nn.LSTM(embed, hidden, layers, batch_first=True)
        self.linear = nn.Linear(hidden, vocab)
        with torch.no_grad():  # models.py:41-45
            self.embed.weight.uniform_(-0.1, 0.1)
            self.linear.weight.uniform_(-0.1, 0.1)
            self.linear.bias.zero_()

    def forward(self, features, captions, lengths):  # models.py:47-54
        steps = torch.cat([features[:, None, :], self.embed(captions)], dim=1)
        packed = pack_padded_sequence(steps, lengths, batch_first=True)
        out, _ = self.lstm(packed)
        return self.linear(out.data)

    @torch.no_grad()
    def sample(self, features, states=None, steps=20):  # models.py:56-67
        x = features[:, None, :]
        picked = []
        for _ in range(steps):
            h, states = self.lstm(x, states)
            tok = self.linear(h[:, 0, :]).max(1, keepdim=True)[1]
            picked.append(tok)
            x = self.embed(tok)
        return torch.cat(picked, 1)


class EncoderHeadCPU(nn.Module):
    """The trainable head of EncoderCNN on pooled ResNet features (models.py:16-17, 22-23, 27-28)."""

    def __init__(self, embed, pooled_dim=2048):
        super().__init__()
        self.fc = nn.Linear(pooled_dim, embed)
        self.bn = nn.BatchNorm1d(embed, momentum=0.01)
        with torch.no_grad():
            self.fc.weight.normal_(0.0, 0.02)
            self.fc.bias.zero_()

    def forward(self, pooled):
        return self.bn(self.fc(pooled))


def train_step(dec, features, captions, lengths, targets):
    """train.py:137-144: zero_grad, forward, CrossEntropyLoss (mean), backward."""
    dec.zero_grad(set_to_none=True)
    loss = nn.functional.cross_entropy(dec(features, captions, lengths), targets)
    loss.backward()
    return loss


def full_step(head, dec, opt, pooled, captions, lengths, targets, grad_clip=0.1):
    """train.py:137-146 for the models.py pair on precomputed pooled features: zero_grad, head + decoder
    forward, CE, backward, clip_gradient (clamp, train.py:88-91), Adam step."""
    opt.zero_grad(set_to_none=True)
    loss = nn.functional.cross_entropy(dec(head(pooled), captions, lengths), targets)
    loss.backward()
    for group in opt.param_groups:
        for p in group["params"]:
            if p.grad is not None:
                p.grad.data.clamp_(-grad_clip, grad_clip)
    opt.step()
    return loss


def time_full_train(B, E, H, V, L, batch, steps=3, warmup=1, threads=None):
    """-> (captions/s, seconds per step, loss) of full_step on batch dict(pooled, captions, lengths, targets)."""
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    head, dec = EncoderHeadCPU(E, batch["pooled"].shape[1]), CaptionDecoderCPU(E, H, V, L)
    opt = torch.optim.Adam(list(head.parameters()) + list(dec.parameters()), lr=1e-3)
    p = torch.from_numpy(batch["pooled"])
    c = torch.from_numpy(batch["captions"])
    t = torch.from_numpy(batch["targets"])
    ts, loss = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = full_step(head, dec, opt, p, c, batch["lengths"], t)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return B / dt, dt, float(loss.detach())


def time_train(B, E, H, V, L, batch, steps=3, warmup=1, threads=None):
    """-> (captions/s, seconds per step, loss)."""
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    dec = CaptionDecoderCPU(E, H, V, L)
    f = torch.from_numpy(batch["features"])
    c = torch.from_numpy(batch["captions"])
    t = torch.from_numpy(batch["targets"])
    for _ in range(warmup):
        train_step(dec, f, c, batch["lengths"], t)
    ts = []
    loss = None
    for _ in range(steps):
        t0 = time.perf_counter()
        loss = train_step(dec, f, c, batch["lengths"], t)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return B / dt, dt, float(loss.detach())


def time_greedy(B, E, H, V, L, features, steps=2, warmup=1, threads=None):
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    dec = CaptionDecoderCPU(E, H, V, L).eval()
    f = torch.from_numpy(features)
    for _ in range(warmup):
        dec.sample(f)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        dec.sample(f)
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    return B * 20 / dt, dt
