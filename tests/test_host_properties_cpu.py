"""Property tests (hypothesis) of the host-side geometry the path relies on: strided sharding of a length-sorted batch
(SURVEY.md §8(e)), packing (models.py:51, train.py:135) and the synthetic COCO-shaped generator.  CPU."""
import numpy as np
from hypothesis import given, settings, strategies as st

import show_and_tell_b200 as snt
from show_and_tell_b200 import parallel
from oracle import snt_oracle as O


@settings(max_examples=60, deadline=None)
@given(B=st.integers(1, 97), world=st.sampled_from([1, 2, 3, 4, 8]), seed=st.integers(0, 10 ** 6))
def test_strided_shards_partition_the_sorted_batch(B, world, seed):
    rng = np.random.default_rng(seed)
    lengths = np.sort(rng.integers(1, 21, size=B))[::-1]
    T = int(lengths[0])
    batch = {"lengths": [int(x) for x in lengths], "captions": rng.integers(0, 50, size=(B, T)),
             "features": rng.standard_normal((B, 8))}
    seen, tokens = [], 0
    for r in range(world):
        sh = parallel.shard_batch(batch, world, r)
        ls = sh["lengths"]
        assert ls == sorted(ls, reverse=True)                     # every shard stays sorted: packable as it is
        assert sh["n_tokens_global"] == int(lengths.sum())
        assert sh["captions"].shape[0] == len(ls) == sh["features"].shape[0]
        tokens += sum(ls)
        for i, row in enumerate(sh["captions"]):
            seen.append((ls[i], tuple(row.tolist())))
    assert tokens == int(lengths.sum())                           # the shards' token counts add up to the global count
    want = sorted((int(lengths[i]), tuple(batch["captions"][i].tolist())) for i in range(B))
    assert sorted(seen) == want                                   # a partition: every row exactly once
    sizes = [len(parallel.shard_batch(batch, world, r)["lengths"]) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1                           # balanced to within one row


@settings(max_examples=60, deadline=None)
@given(B=st.integers(1, 40), seed=st.integers(0, 10 ** 6))
def test_pack_host_is_pack_padded_sequence(B, seed):
    from torch.nn.utils.rnn import pack_padded_sequence
    import torch
    rng = np.random.default_rng(seed)
    lengths = np.sort(rng.integers(1, 21, size=B))[::-1].copy()
    caps = rng.integers(0, 1000, size=(B, int(lengths[0]))).astype(np.int64)
    ours = snt.synthetic.pack_host(caps, [int(x) for x in lengths])
    ref = pack_padded_sequence(torch.from_numpy(caps), [int(x) for x in lengths], batch_first=True)   # train.py:135
    np.testing.assert_array_equal(ours, ref[0].numpy())
    np.testing.assert_array_equal(O.pack_rows(caps, [int(x) for x in lengths]), ref[0].numpy())
    T, bs, off = O.pack_info([int(x) for x in lengths])
    np.testing.assert_array_equal(bs, ref.batch_sizes.numpy())
    assert off[-1] == int(lengths.sum()) and T == int(lengths[0])


@settings(max_examples=20, deadline=None)
@given(B=st.integers(1, 300), V=st.integers(10, 5000), seed=st.integers(0, 1000))
def test_synthetic_batches_are_coco_shaped(B, V, seed):
    b = snt.synthetic.make_batch(B, V, embed=16, seed=seed)
    ls = np.asarray(b["lengths"])
    assert (np.diff(ls) <= 0).all() and ls.min() >= 1 and ls.max() <= 21       # sorted descending (data_loader.py:50)
    caps = b["captions"]
    assert caps.dtype == np.int64 and caps.shape[0] == B and caps.min() >= 0 and caps.max() < V
    assert b["features"].shape == (B, 16) and b["features"].dtype == np.float32
