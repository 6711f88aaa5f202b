"""The numpy oracle against the reference ITSELF (the unmodified models.py that build() places under oracle/_ref) on seeded
random cases beyond the committed goldens: ragged and equal lengths, one and two layers, batch 1, captions of width 1.
fp64 on both sides, so the bars are rounding-level.  CPU only; skipped where oracle/_ref cannot be populated (it can in
the build container and it travels to the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_arm as RA
from oracle import snt_oracle as O

CASES = [  # B, E, H, V, L, width, ragged
    (7, 16, 24, 50, 1, 9, True),
    (12, 32, 32, 97, 2, 14, True),
    (1, 8, 16, 30, 1, 5, False),
    (5, 16, 16, 40, 2, 1, False),
    (9, 24, 40, 120, 1, 20, True),
    (16, 8, 8, 11, 3, 6, True),
]


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_case(B, E, V, width, ragged, seed):
    rng = np.random.default_rng(seed)
    if ragged:
        lengths = np.sort(rng.integers(1, width + 2, size=B))[::-1].copy()
        lengths[0] = width + 1                      # the widest caption defines the padded width
    else:
        lengths = np.full(B, width + 1)
    captions = rng.integers(0, V, size=(B, width)).astype(np.int64)
    features = rng.standard_normal((B, E))
    return features, captions, [int(x) for x in lengths]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_E%d_H%d_V%d_L%d_w%d_%s" % (c[:6] + ("ragged" if c[6] else "equal",)))
def test_oracle_matches_reference_module(case):
    if not RA.available() and RA.build_ref() is None:
        pytest.skip("oracle/_ref is only populated where /root/reference exists (build container)")
    B, E, H, V, L, width, ragged = case
    ref = RA.load()
    torch.manual_seed(1000 + B + 7 * L)
    dec = ref.DecoderRNN(E, H, V, L).double()
    with torch.no_grad():                            # the reference's init leaves the LSTM at torch's default; widen it
        for p in dec.lstm.parameters():
            p.mul_(3.0)
    features, captions, lengths = make_case(B, E, V, width, ragged, seed=B * 31 + V)
    params = {k: v.detach().numpy().copy() for k, v in dec.state_dict().items()}
    targets = O.pack_rows(np.concatenate([captions, np.zeros((B, 1), np.int64)], 1), lengths)  # any ids of the right shape
    targets = (targets + 3) % V

    f = torch.from_numpy(features).requires_grad_(True)
    dec.zero_grad()
    logits = dec(f, torch.from_numpy(captions), lengths)              # models.py:47-54
    loss = torch.nn.CrossEntropyLoss()(logits, torch.from_numpy(targets))
    loss.backward()

    r = O.train_step(params, features, captions, lengths, targets)
    assert rel(r["logits"], logits.detach().numpy()) < 1e-12
    assert abs(r["loss"] - float(loss.detach())) / float(loss.detach()) < 1e-12
    assert rel(r["dfeatures"], f.grad.numpy()) < 1e-10
    for k, p in dec.named_parameters():
        assert rel(r["grads"][k], p.grad.numpy()) < 1e-10, k

    ids_ref = RA.sample_keepdim(dec.eval(), torch.from_numpy(features)).numpy()   # models.py:56-67
    ids, margins = O.greedy_sample(params, features, return_margins=True)
    bad = (ids != ids_ref) & (margins > 1e-9)
    first_bad = [int(np.argmax(row)) if row.any() else -1 for row in (ids != ids_ref)]
    # after a (margin-gated) divergence the two loops follow different tokens: compare up to the first mismatch only
    for b in range(B):
        k = first_bad[b]
        assert k < 0 or margins[b, k] <= 1e-9, (b, k, margins[b, k])
    assert not bad[:, 0].any()
