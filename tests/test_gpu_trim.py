"""f4: the caller-side tail of sample() — eval.py:101-109 keeps the words before the first
<end> — on the device, through the C ABI (snt_caption_trim), bit-exact against the CPU oracle's restatement of the
reference's loop."""
import numpy as np
import pytest
import torch

from conftest import decoder_from_params, golden_params, load_golden
from oracle import snt_oracle as O

pytestmark = pytest.mark.gpu


def _ids(B, S, V, p_end, seed):
    rng = np.random.default_rng(seed)
    ids = rng.integers(3, V, size=(B, S)).astype(np.int64)
    ids[rng.random((B, S)) < p_end] = 2
    return ids


@pytest.mark.parametrize("B,S,p_end", [(1, 20, 0.1), (7, 20, 0.0), (33, 20, 0.08), (4096, 20, 0.05), (5, 1, 0.5),
                                      (9, 32, 0.03), (8, 33, 0.03), (3, 100, 0.01), (17, 64, 1.0)])
def test_trim_matches_oracle(B, S, p_end):
    import show_and_tell_b200 as snt
    ids = _ids(B, S, 1000, p_end, seed=B * 1000 + S)
    ref_ids, ref_len = O.trim_captions(ids)
    out, lengths = snt.ops.trim_captions(torch.from_numpy(ids).cuda())
    assert lengths.dtype == torch.int32 and out.dtype == torch.int64 and tuple(out.shape) == (B, S)
    assert np.array_equal(lengths.cpu().numpy(), ref_len)
    assert np.array_equal(out.cpu().numpy(), ref_ids)
    # lengths only; other end / pad ids; in place through the raw entry point
    none, l2 = snt.ops.trim_captions(torch.from_numpy(ids).cuda(), end_id=7, pad_id=-1, return_ids=False)
    assert none is None and np.array_equal(l2.cpu().numpy(), O.trim_captions(ids, 7, -1)[1])
    t = torch.from_numpy(ids).cuda()
    L = snt._lib
    L.call("snt_caption_trim", L.ptr(t), B, S, 2, 0, None, L.ptr(t), L.stream_ptr())
    assert np.array_equal(t.cpu().numpy(), ref_ids)


def test_trim_edge_rows():
    import show_and_tell_b200 as snt
    ids = np.array([[2] + [5] * 19,                 # <end> first: empty caption
                    [5] * 19 + [2],                 # <end> last
                    [5] * 20,                       # no <end>: all 20 words kept
                    [5, 2, 6, 2] + [2] * 16], dtype=np.int64)
    out, lengths = snt.ops.trim_captions(torch.from_numpy(ids).cuda())
    assert lengths.cpu().tolist() == [0, 19, 20, 1]
    assert np.array_equal(out.cpu().numpy(), O.trim_captions(ids)[0])
    one, l1 = snt.ops.trim_captions(torch.from_numpy(ids[3]).cuda())       # sample() squeezes a batch of one
    assert tuple(one.shape) == (20,) and l1.cpu().tolist() == [1]
    with pytest.raises(RuntimeError):
        snt.ops.trim_captions(torch.zeros(2, 20, dtype=torch.int32, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        snt.ops.trim_captions(torch.zeros(2, 20, dtype=torch.int64))


def test_sample_trimmed_equals_sample_then_reference_loop():
    g = load_golden("dec_l1_mid")
    dec = decoder_from_params(golden_params(g), "fp32").eval()
    feats = torch.from_numpy(g["features"]).cuda()
    ids = dec.sample(feats)
    # random-init greedy rows rarely contain <end>: plant one at position b % 21 (none when that is 20)
    pos = torch.arange(ids.shape[0], device="cuda") % 21
    planted = torch.where(torch.arange(20, device="cuda")[None, :] == pos[:, None], torch.full_like(ids, 2), ids)
    out, lengths = dec.sample_trimmed(feats)
    ref_ids, ref_len = O.trim_captions(ids.cpu().numpy())
    assert np.array_equal(out.cpu().numpy(), ref_ids) and np.array_equal(lengths.cpu().numpy(), ref_len)
    import show_and_tell_b200 as snt
    out2, len2 = snt.ops.trim_captions(planted)
    ref2 = O.trim_captions(planted.cpu().numpy())
    assert np.array_equal(out2.cpu().numpy(), ref2[0]) and np.array_equal(len2.cpu().numpy(), ref2[1])
    assert int(len2.min()) == 0 and int(len2.max()) == 20
