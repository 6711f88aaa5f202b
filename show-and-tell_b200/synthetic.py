"""Synthetic COCO-shaped batches for the caption-decoder path.

The shapes follow what the reference's loader hands the model: `collate_fn`
(/root/reference/data_loader.py:48-62) sorts by caption length (descending) and zero-pads;
`preprocess.py:75-78` fixes <pad>=0, <start>=1, <end>=2, <unk>=3.  Pure numpy on the host — this is
input generation, not part of the measured path.
"""
from __future__ import annotations

import numpy as np

PAD, START, END, UNK = 0, 1, 2, 3


def make_lengths(batch, rng, mean=12.5, std=2.5, lo=6, hi=20):
    l = np.clip(np.rint(rng.normal(mean, std, size=batch)), lo, hi).astype(np.int64)
    return np.sort(l)[::-1].copy()


def make_batch(batch, vocab, embed=None, seed=1, lengths=None, pooled_dim=None):
    """Returns dict(captions[B,Tmax] i64, lengths list[int] desc, features[B,E] f32 (if embed),
    pooled[B,pooled_dim] f32 (if pooled_dim))."""
    rng = np.random.default_rng(seed)
    if lengths is None:
        lengths = make_lengths(batch, rng)
    lengths = np.asarray(lengths, dtype=np.int64)
    tmax = int(lengths.max())
    caps = np.zeros((batch, tmax), dtype=np.int64)
    body = rng.integers(4, vocab, size=(batch, tmax), dtype=np.int64) if vocab > 4 else \
        np.full((batch, tmax), UNK, dtype=np.int64)
    for i, l in enumerate(lengths):
        l = int(l)
        caps[i, :l] = body[i, :l]
        caps[i, 0] = START
        caps[i, l - 1] = END
    out = dict(captions=caps, lengths=[int(x) for x in lengths])
    if embed is not None:
        out["features"] = rng.standard_normal((batch, embed)).astype(np.float32)
    if pooled_dim is not None:
        out["pooled"] = (0.5 * np.abs(rng.standard_normal((batch, pooled_dim)))).astype(np.float32)
    return out


def pack_host(padded, lengths):
    """Host-side pack_padded_sequence(padded, lengths, batch_first=True)[0] (time-major rows)."""
    la = np.asarray(lengths)
    rows = [padded[: int((la > t).sum()), t] for t in range(int(la[0]))]
    return np.concatenate(rows, 0)


def shard_rows(batch, world, rank):
    """Strided batch sharding row i -> rank i % world: every shard of a length-sorted batch stays
    sorted and token counts stay balanced."""
    return np.arange(rank, batch, world)
