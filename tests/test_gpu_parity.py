"""Parity of the CUDA path (through the C ABI, via the reference-shaped modules) against
 (1) the committed golden vectors produced by the unmodified reference (tests/golden, oracle/make_golden.py),
 (2) the CPU oracle (oracle/snt_oracle.py) on seeded inputs at sizes it finishes in seconds,
 (3) size-independent properties at BASELINE.json's full sizes.
Tolerances (north_star): fp32 mode loss/logits 1e-5 / grads 1e-4 rel; bf16 mode loss 1e-3 rel, grads 1e-2 rel;
greedy tokens exact in fp32 mode wherever the fp64 top-2 margin exceeds rounding noise (SURVEY.md §7 hard part 1).
"""
import numpy as np
import pytest
import torch

from conftest import decoder_from_params, golden_params, load_golden, rel_err
from oracle import snt_oracle as O

pytestmark = pytest.mark.gpu

DEC_CASES = ["dec_l1_a", "dec_l2_b", "dec_b1", "dec_l1_mid"]
BF16_CASES = ["dec_l1_a", "dec_b1", "dec_l1_mid"]  # E and H multiples of 8 (TMA pitch rule)
TOL = {"fp32": dict(loss=1e-5, logits=2e-5, grad=1e-4), "bf16": dict(loss=1e-3, logits=1e-2, grad=1e-2)}


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _grads(dec):
    return {k: p.grad.detach().cpu().numpy() for k, p in dec.named_parameters()}


def _run_loss(dec, feats, caps, lengths, targets):
    dec.zero_grad(set_to_none=True)
    f = feats.clone().requires_grad_(True)
    loss = dec.loss(f, caps, lengths, targets)
    loss.backward()
    return float(loss), _grads(dec), f.grad.cpu().numpy()


def _run_logits(dec, feats, caps, lengths, targets):
    dec.zero_grad(set_to_none=True)
    f = feats.clone().requires_grad_(True)
    logits = dec(f, caps, lengths)                                # models.py:47-54
    loss = torch.nn.CrossEntropyLoss()(logits, targets)           # train.py:53,143 — torch's CE on OUR logits
    loss.backward()
    return logits.detach().cpu().numpy(), float(loss), _grads(dec), f.grad.cpu().numpy()


@pytest.mark.parametrize("name", DEC_CASES)
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_train_step_vs_reference_golden(name, prec):
    if prec == "bf16" and name not in BF16_CASES:
        pytest.skip("bf16 mode needs E and H to be multiples of 8")
    g = load_golden(name)
    tol = TOL[prec]
    dec = decoder_from_params(golden_params(g), prec)
    feats, caps, targets = _t(g["features"]), _t(g["captions"]), _t(g["targets"])
    lengths = g["lengths"].tolist()
    # fused path
    loss, grads, dfeat = _run_loss(dec, feats, caps, lengths, targets)
    assert abs(loss - g["f64.loss"]) / g["f64.loss"] < tol["loss"]
    for k, v in grads.items():
        assert rel_err(v, g[f"f64.grad.{k}"]) < tol["grad"], (k, rel_err(v, g[f"f64.grad.{k}"]))
    assert rel_err(dfeat, g["f64.dfeatures"]) < tol["grad"]
    # strict drop-in path: materialised logits + torch's own CrossEntropyLoss
    logits, loss2, grads2, dfeat2 = _run_logits(dec, feats, caps, lengths, targets)
    assert logits.shape == g["f64.logits"].shape
    assert rel_err(logits, g["f64.logits"]) < tol["logits"]
    assert abs(loss2 - g["f64.loss"]) / g["f64.loss"] < tol["loss"]
    for k, v in grads2.items():
        assert rel_err(v, g[f"f64.grad.{k}"]) < tol["grad"], (k, rel_err(v, g[f"f64.grad.{k}"]))
    assert rel_err(dfeat2, g["f64.dfeatures"]) < tol["grad"]


@pytest.mark.parametrize("name", DEC_CASES)
def test_greedy_fp32_vs_reference_golden(name):
    g = load_golden(name)
    dec = decoder_from_params(golden_params(g), "fp32").eval()
    ids = dec.sample(_t(g["features"]))
    B = int(g["B"])
    assert ids.dtype == torch.int64
    assert tuple(ids.shape) == ((20,) if B == 1 else (B, 20))      # models.py:67 .squeeze()
    ids = ids.reshape(B, 20).cpu().numpy()
    ref = g["f64.greedy_ids"]
    marg = g["f64.greedy_margins"]
    # exact until (and unless) a row meets a step whose fp64 margin is within fp32 rounding noise
    for b in range(B):
        for s in range(20):
            if ids[b, s] != ref[b, s]:
                assert marg[b, s] < 1e-5, (b, s, ids[b, s], ref[b, s], marg[b, s])
                break
    assert (ids == ref).mean() > 0.95


@pytest.mark.parametrize("name", BF16_CASES)
def test_greedy_bf16_vs_reference_golden(name):
    g = load_golden(name)
    dec = decoder_from_params(golden_params(g), "bf16").eval()
    B = int(g["B"])
    ids = dec.sample(_t(g["features"]), precision="bf16").reshape(B, 20).cpu().numpy()
    ref, marg = g["f64.greedy_ids"], g["f64.greedy_margins"]
    for b in range(B):
        for s in range(20):
            if ids[b, s] != ref[b, s]:
                assert marg[b, s] < 3e-2, (b, s, marg[b, s])   # bf16 operand rounding noise on the logits
                break


def test_greedy_states_and_steps():
    g = load_golden("dec_l2_b")
    p = golden_params(g)
    dec = decoder_from_params(p, "fp32").eval()
    B, H, L = int(g["B"]), int(g["H"]), int(g["L"])
    rng = np.random.default_rng(0)
    h0 = rng.standard_normal((L, B, H)).astype(np.float32) * 0.1
    c0 = rng.standard_normal((L, B, H)).astype(np.float32) * 0.1
    ids = dec.sample(_t(g["features"]), (_t(h0), _t(c0))).cpu().numpy()
    ref, marg = O.greedy_sample(O.cast_params(p, np.float64), g["features"].astype(np.float64),
                                states=(h0.astype(np.float64), c0.astype(np.float64)), return_margins=True)
    for b in range(B):
        for s in range(20):
            if ids[b, s] != ref[b, s]:
                assert marg[b, s] < 1e-5
                break
    with pytest.raises(RuntimeError):
        dec.sample(_t(g["features"]), (_t(h0[:, :2]), _t(c0[:, :2])))


@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_head_vs_reference_golden(mode, prec):
    import show_and_tell_b200 as snt
    g = load_golden("head")
    hp = golden_params(g)
    E = int(g["E"])
    enc = snt.EncoderCNN(E, backbone=False, precision=prec)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in hp.items()}
    if mode == "eval":
        sd["bn.running_mean"] = torch.from_numpy(g["eval.running_mean_before"])
        sd["bn.running_var"] = torch.from_numpy(g["eval.running_var_before"])
    enc.load_state_dict(sd, strict=False)
    enc = enc.cuda()
    enc.train(mode == "train")
    out = enc.forward_pooled(_t(g["pooled"]))
    out.backward(_t(g["dout"]))
    ft, gt = (2e-5, 1e-4) if prec == "fp32" else (1e-2, 2e-2)
    assert rel_err(out.detach().cpu().numpy(), g[mode + ".features"]) < ft
    assert rel_err(enc.bn.running_mean.cpu().numpy(), g[mode + ".running_mean_after"]) < ft
    assert rel_err(enc.bn.running_var.cpu().numpy(), g[mode + ".running_var_after"]) < ft
    dn = np.linalg.norm(g["dout"])
    for k, p in enc.named_parameters():
        ref = g[f"{mode}.grad.{k}"]
        # train-mode fc.bias grad is analytically zero (BN removes the batch mean): absolute floor
        assert np.linalg.norm(p.grad.cpu().numpy() - ref) < gt * max(np.linalg.norm(ref), dn), k


def test_clamp_adam_vs_reference_golden():
    import show_and_tell_b200 as snt
    g = load_golden("adam")
    p = _t(g["p0"].astype(np.float32))
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for s in range(g["grads"].shape[0]):
        snt.ops.clamp_adam_(p, _t(g["grads"][s].astype(np.float32)), m, v, s + 1)
        np.testing.assert_allclose(p.cpu().numpy(), g["params"][s], rtol=3e-6, atol=3e-7)
    np.testing.assert_allclose(m.cpu().numpy(), g["exp_avg"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(v.cpu().numpy(), g["exp_avg_sq"], rtol=1e-5, atol=1e-10)


def test_clamp_adam_multi_matches_single():
    import show_and_tell_b200 as snt
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(7,), (1000, 33), (5, 5), (4096,), (3,)] * 6          # 30 tensors: more than one table
    ps = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    gs = [torch.randn(s, device="cuda", generator=g) * 0.3 for s in shapes]
    ref = [(p.clone(), torch.zeros_like(p), torch.zeros_like(p)) for p in ps]
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    for step in (1, 2, 3):
        snt.ops.clamp_adam_multi_(ps, gs, ms, vs, step)
        for (rp, rm, rv), gg in zip(ref, gs):
            snt.ops.clamp_adam_(rp, gg, rm, rv, step)
    for p, m, v, (rp, rm, rv) in zip(ps, ms, vs, ref):
        assert torch.equal(p, rp) and torch.equal(m, rm) and torch.equal(v, rv)


# ---------------------------------------------------------------------------------------------------------
# seeded mid-size cases against the CPU oracle (fp64)
# ---------------------------------------------------------------------------------------------------------
def _oracle_case(B, E, H, V, L, seed, lengths=None):
    import show_and_tell_b200 as snt
    torch.manual_seed(seed)
    dec = snt.DecoderRNN(E, H, V, L)
    params = {k: v.detach().numpy().copy() for k, v in dec.state_dict().items()}
    b = snt.synthetic.make_batch(B, V, embed=E, seed=seed + 1, lengths=lengths)
    targets = snt.synthetic.pack_host(b["captions"], b["lengths"])
    return params, b, targets


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("B,E,H,V,L", [(64, 64, 128, 1000, 1), (48, 32, 64, 517, 2), (200, 256, 512, 2000, 1),
                                         (150, 64, 256, 600, 2), (300, 128, 512, 1200, 2)])
def test_train_step_vs_oracle(prec, B, E, H, V, L):
    params, b, targets = _oracle_case(B, E, H, V, L, seed=B + L)
    r = O.train_step(O.cast_params(params, np.float64), b["features"].astype(np.float64), b["captions"],
                     b["lengths"], targets)
    dec = decoder_from_params(params, prec)
    loss, grads, dfeat = _run_loss(dec, _t(b["features"]), _t(b["captions"]), b["lengths"], _t(targets))
    tol = TOL[prec]
    assert abs(loss - r["loss"]) / r["loss"] < tol["loss"]
    for k, v in grads.items():
        assert rel_err(v, r["grads"][k]) < tol["grad"], (k, rel_err(v, r["grads"][k]))
    assert rel_err(dfeat, r["dfeatures"]) < tol["grad"]


@pytest.mark.parametrize("lengths", [[5] * 9, [20, 1, 1, 1], [3, 2, 1], [1], [1, 1, 1]])
def test_edge_length_patterns_vs_oracle(lengths):
    B = len(lengths)
    params, b, targets = _oracle_case(B, 16, 24, 50, 1, seed=3, lengths=lengths)
    r = O.train_step(O.cast_params(params, np.float64), b["features"].astype(np.float64), b["captions"],
                     b["lengths"], targets)
    dec = decoder_from_params(params, "fp32")
    loss, grads, dfeat = _run_loss(dec, _t(b["features"]), _t(b["captions"]), b["lengths"], _t(targets))
    assert abs(loss - r["loss"]) / r["loss"] < 1e-5
    for k, v in grads.items():
        assert np.linalg.norm(v - r["grads"][k]) < 1e-4 * max(np.linalg.norm(r["grads"][k]), 1e-3), k
    assert rel_err(dfeat, r["dfeatures"]) < 1e-4


def test_convention_b_and_wide_captions():
    """train.py:134-139 convention: captions[:, :-1] with lengths-1; and Tc+1 > max(lengths)."""
    params, b, _ = _oracle_case(12, 16, 24, 60, 1, seed=5)
    caps, lengths = b["captions"], b["lengths"]
    l1 = [l - 1 for l in lengths]
    targets = O.pack_rows(caps[:, 1:], l1)
    r = O.train_step(O.cast_params(params, np.float64), b["features"].astype(np.float64), caps[:, :-1], l1, targets)
    dec = decoder_from_params(params, "fp32")
    loss, grads, _ = _run_loss(dec, _t(b["features"]), _t(caps[:, :-1]), l1, _t(targets))
    assert abs(loss - r["loss"]) / r["loss"] < 1e-5
    wide = np.concatenate([caps, np.zeros((caps.shape[0], 5), np.int64)], 1)   # extra zero padding columns
    t2 = O.pack_rows(caps, lengths)
    r2 = O.train_step(O.cast_params(params, np.float64), b["features"].astype(np.float64), caps, lengths, t2)
    loss2, _, _ = _run_loss(dec, _t(b["features"]), _t(wide), lengths, _t(t2))
    assert abs(loss2 - r2["loss"]) / r2["loss"] < 1e-5


def test_errors_fail_loudly():
    import show_and_tell_b200 as snt
    params, b, targets = _oracle_case(6, 16, 24, 60, 1, seed=7)
    dec = decoder_from_params(params, "fp32")
    f, c = _t(b["features"]), _t(b["captions"])
    with pytest.raises(RuntimeError):
        dec(f, c, [1, 2, 3, 4, 5, 6])                                    # not sorted descending
    with pytest.raises(RuntimeError):
        dec(f, c, [c.shape[1] + 2] + b["lengths"][1:])                  # longer than captions allow
    with pytest.raises(RuntimeError):
        dec(f, c, b["lengths"][:-1])                                     # batch mismatch
    with pytest.raises(RuntimeError):
        dec.loss(f, c, b["lengths"], _t(targets)[:-1])                   # target count mismatch
    with pytest.raises(RuntimeError):
        snt.DecoderRNN(16, 24, 60, 1).loss(torch.from_numpy(b["features"]), torch.from_numpy(b["captions"]),
                                           b["lengths"], torch.from_numpy(targets))  # CPU tensors: no fallback
    bad = c.clone()
    bad[0, 0] = 60                                                       # token id == V
    dec(f, bad, b["lengths"])
    assert snt._lib.read_flags() & 1
    assert snt._lib.read_flags() == 0


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[1] at full size: size-independent properties
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_full_size_properties(prec):
    import show_and_tell_b200 as snt
    B, E, H, V = 1024, 256, 512, 10000
    torch.manual_seed(0)
    dec = snt.DecoderRNN(E, H, V, 1, precision=prec).cuda()
    b = snt.synthetic.make_batch(B, V, embed=E, seed=1)
    targets = _t(snt.synthetic.pack_host(b["captions"], b["lengths"]))
    feats, caps = _t(b["features"]), _t(b["captions"])
    loss1, g1, df1 = _run_loss(dec, feats, caps, b["lengths"], targets)
    loss2, g2, df2 = _run_loss(dec, feats, caps, b["lengths"], targets)
    assert abs(loss1 - np.log(V)) < 0.05                      # random init: loss ~ ln V
    assert loss1 == loss2                                      # run-to-run determinism
    for k in g1:
        assert np.array_equal(g1[k], g2[k]), k
    assert np.array_equal(g1["lstm.bias_ih_l0"], g1["lstm.bias_hh_l0"])
    # softmax - onehot sums to zero over the vocabulary => the vocab bias gradient sums to ~0
    assert abs(g1["linear.bias"].sum()) < 1e-3 * np.abs(g1["linear.bias"]).sum()
    used = np.zeros(V, bool)
    used[np.unique(b["captions"])] = True                      # tokens that occur as an input
    assert np.all(g1["embed.weight"][~used] == 0)
    assert np.isfinite(df1).all() and np.abs(df1).sum() > 0
    # fused loss == torch CE on materialised logits
    logits, loss3, g3, _ = _run_logits(dec, feats, caps, b["lengths"], targets)
    assert logits.shape == (int(sum(b["lengths"])), V)
    assert abs(loss3 - loss1) / loss1 < (1e-6 if prec == "fp32" else 1e-4)
    assert rel_err(g3["linear.weight"], g1["linear.weight"]) < (1e-5 if prec == "fp32" else 1e-2)


def test_full_size_bf16_matches_fp32_mode():
    import show_and_tell_b200 as snt
    B, E, H, V = 1024, 256, 512, 10000
    torch.manual_seed(0)
    dec = snt.DecoderRNN(E, H, V, 1, precision="fp32").cuda()
    b = snt.synthetic.make_batch(B, V, embed=E, seed=1)
    targets = _t(snt.synthetic.pack_host(b["captions"], b["lengths"]))
    feats, caps = _t(b["features"]), _t(b["captions"])
    l32, g32, d32 = _run_loss(dec, feats, caps, b["lengths"], targets)
    dec.precision = "bf16"
    l16, g16, d16 = _run_loss(dec, feats, caps, b["lengths"], targets)
    assert abs(l16 - l32) / l32 < 1e-3
    for k in g32:
        assert rel_err(g16[k], g32[k]) < 1e-2, (k, rel_err(g16[k], g32[k]))
    assert rel_err(d16, d32) < 1e-2


def test_scaled_config_bf16_matches_fp32_mode():
    """BASELINE.json configs[3]: embed 512, hidden 1024, 2-layer LSTM, 32k vocab, batch 2048 (fused vocab CE)."""
    import show_and_tell_b200 as snt
    B, E, H, V, L = 2048, 512, 1024, 32000, 2
    torch.manual_seed(0)
    dec = snt.DecoderRNN(E, H, V, L, precision="fp32").cuda()
    b = snt.synthetic.make_batch(B, V, embed=E, seed=1)
    targets = _t(snt.synthetic.pack_host(b["captions"], b["lengths"]))
    feats, caps = _t(b["features"]), _t(b["captions"])
    l32, g32, d32 = _run_loss(dec, feats, caps, b["lengths"], targets)
    assert abs(l32 - np.log(V)) < 0.05
    dec.precision = "bf16"
    l16, g16, d16 = _run_loss(dec, feats, caps, b["lengths"], targets)
    l16b, g16b, _ = _run_loss(dec, feats, caps, b["lengths"], targets)
    assert l16 == l16b and all(np.array_equal(g16[k], g16b[k]) for k in g16)      # deterministic
    assert abs(l16 - l32) / l32 < 1e-3
    for k in g32:
        assert rel_err(g16[k], g32[k]) < 1e-2, (k, rel_err(g16[k], g32[k]))
    assert rel_err(d16, d32) < 1e-2


def test_full_size_greedy_fp32_vs_bf16_first_token():
    import show_and_tell_b200 as snt
    B, E, H, V = 4096, 256, 512, 10000
    torch.manual_seed(0)
    dec = snt.DecoderRNN(E, H, V, 1).cuda().eval()
    feats = _t(snt.synthetic.make_batch(B, V, embed=E, seed=1)["features"])
    a = dec.sample(feats, precision="fp32")
    b = dec.sample(feats, precision="fp32")
    assert tuple(a.shape) == (B, 20) and a.dtype == torch.int64
    assert torch.equal(a, b)                                   # deterministic
    assert int(a.min()) >= 0 and int(a.max()) < V
    c = dec.sample(feats, precision="bf16")
    assert (a[:, 0] == c[:, 0]).float().mean() > 0.9           # first token agrees except near-ties
