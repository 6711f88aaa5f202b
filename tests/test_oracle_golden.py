"""The oracle (numpy restatement) against vectors produced by the unmodified reference
(oracle/make_golden.py imports /root/reference/models.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_params, load_golden
from oracle import snt_oracle as O

DEC_CASES = ["dec_l1_a", "dec_l2_b", "dec_b1", "dec_l1_mid"]


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) /
                 max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))


@pytest.mark.parametrize("name", DEC_CASES)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_decoder_train_step_matches_reference(name, prec):
    g = load_golden(name)
    dt = np.float32 if prec == "f32" else np.float64
    tol = 2e-5 if prec == "f32" else 1e-11
    p = O.cast_params(golden_params(g), dt)
    r = O.train_step(p, g["features"].astype(dt), g["captions"], g["lengths"].tolist(), g["targets"])
    assert rel(r["logits"], g[prec + ".logits"]) < tol
    assert abs(r["loss"] - g[prec + ".loss"]) / g[prec + ".loss"] < tol
    assert rel(r["dfeatures"], g[prec + ".dfeatures"]) < tol * 5
    for k, v in r["grads"].items():
        assert rel(v, g[f"{prec}.grad.{k}"]) < tol * 5, k


@pytest.mark.parametrize("name", DEC_CASES)
def test_targets_are_packed_captions(name):
    g = load_golden(name)
    if str(g["convention"]) == "a":
        np.testing.assert_array_equal(O.pack_rows(g["captions"], g["lengths"].tolist()), g["targets"])


@pytest.mark.parametrize("name", DEC_CASES)
def test_greedy_matches_reference(name):
    g = load_golden(name)
    p64 = O.cast_params(golden_params(g), np.float64)
    ids, margins = O.greedy_sample(p64, g["features"].astype(np.float64), return_margins=True)
    assert ids.shape == (int(g["B"]), 20) and ids.dtype == np.int64
    np.testing.assert_array_equal(ids, g["f64.greedy_ids"])
    np.testing.assert_allclose(margins, g["f64.greedy_margins"], rtol=1e-7, atol=1e-12)
    # fp32: equal wherever the fp64 margin is not within rounding noise; teacher-force the reference ids
    p32 = O.cast_params(golden_params(g), np.float32)
    ids32 = O.greedy_sample(p32, g["features"], forced_ids=g["f32.greedy_ids"])
    bad = (ids32 != g["f32.greedy_ids"]) & (g["f32.greedy_margins"] > 1e-5)
    assert not bad.any()
    assert int(g["verbatim_sample_runs"]) == 0  # why the keepdim shim exists


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_head_matches_reference(mode):
    g = load_golden("head")
    hp = golden_params(g)
    if mode == "eval":
        hp["bn.running_mean"], hp["bn.running_var"] = g["eval.running_mean_before"], g["eval.running_var_before"]
    out, cache, rm, rv = O.head_forward(hp, g["pooled"], training=(mode == "train"))
    assert rel(out, g[mode + ".features"]) < 2e-5
    assert rel(rm, g[mode + ".running_mean_after"]) < 1e-5
    assert rel(rv, g[mode + ".running_var_after"]) < 1e-5
    gr = O.head_backward(hp, cache, g["dout"])
    for k, v in gr.items():
        ref = g[f"{mode}.grad.{k}"]
        # train-mode fc.bias grad is analytically zero (BN removes the batch mean): absolute floor
        assert np.linalg.norm(v - ref) < 1e-4 * max(np.linalg.norm(ref), np.linalg.norm(g["dout"])), k


def test_clamp_adam_matches_reference():
    g = load_golden("adam")
    p, m, v = g["p0"].astype(np.float64), np.zeros(64), np.zeros(64)
    for s in range(g["grads"].shape[0]):
        p, m, v = O.clamp_adam(p, g["grads"][s].astype(np.float64), m, v, s + 1)
        np.testing.assert_allclose(p, g["params"][s], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(m, g["exp_avg"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(v, g["exp_avg_sq"], rtol=1e-5, atol=1e-10)


def test_pack_info_edges():
    T, bs, off = O.pack_info([3, 3, 1])
    assert T == 3 and bs.tolist() == [3, 2, 2] and off.tolist() == [0, 3, 5, 7]
    with pytest.raises(ValueError):
        O.pack_info([2, 3])
    with pytest.raises(ValueError):
        O.pack_info([2, 0])
    with pytest.raises(ValueError):
        O.pack_info([])


@pytest.mark.parametrize("name", DEC_CASES)
def test_torch_port_matches_reference(name):
    """The torch CPU baseline bench.py times (oracle/torch_port.py) against the reference's own outputs."""
    import torch
    from oracle import torch_port as TP
    g = load_golden(name)
    dec = TP.CaptionDecoderCPU(int(g["E"]), int(g["H"]), int(g["V"]), int(g["L"]))
    dec.load_state_dict({k: torch.from_numpy(v) for k, v in golden_params(g).items()})
    f, c, t = torch.from_numpy(g["features"]), torch.from_numpy(g["captions"]), torch.from_numpy(g["targets"])
    loss = TP.train_step(dec, f, c, g["lengths"].tolist(), t)
    assert abs(float(loss) - g["f32.loss"]) / g["f32.loss"] < 1e-6
    for k, p in dec.named_parameters():
        assert rel(p.grad.numpy(), g[f"f32.grad.{k}"]) < 1e-5, k
    ids = dec.eval().sample(f).numpy()
    np.testing.assert_array_equal(ids, g["f32.greedy_ids"])


def test_trim_captions_follows_eval_loop():
    """eval.py:101-109 with a stand-in vocabulary: join the words before '<end>'."""
    idx2word = {0: "<pad>", 1: "<start>", 2: "<end>", 3: "<unk>", 4: "a", 5: "dog", 6: "runs"}
    ids = np.array([[4, 5, 6, 2, 4, 2], [2, 4, 4, 4, 4, 4], [4, 5, 4, 5, 4, 5], [1, 4, 3, 6, 5, 2]], dtype=np.int64)
    sentences = []
    for sentence_ids in ids:                      # the reference's loop, verbatim in structure
        sampled_caption = []
        for word_id in sentence_ids:
            word = idx2word[int(word_id)]
            if word == "<end>":
                break
            sampled_caption.append(word)
        sentences.append(" ".join(sampled_caption))
    out, lengths = O.trim_captions(ids)
    assert lengths.tolist() == [3, 0, 6, 5]
    for b in range(len(ids)):
        assert " ".join(idx2word[int(w)] for w in out[b, :lengths[b]]) == sentences[b]
        assert (out[b, lengths[b]:] == 0).all()


@pytest.mark.parametrize("name", DEC_CASES)
def test_reference_arm_reproduces_goldens(name):
    """bench.py's `kind: "reference"` arm (oracle/ref_arm.py over the unmodified models.py under oracle/_ref) gives the
    vectors the reference gave when the goldens were made, and the restated port agrees with it."""
    import torch
    from oracle import ref_arm as RA
    if not RA.available() and RA.build_ref() is None:
        pytest.skip("oracle/_ref is only populated where /root/reference exists (build container)")
    ref = RA.load()
    g = load_golden(name)
    dec = ref.DecoderRNN(int(g["E"]), int(g["H"]), int(g["V"]), int(g["L"]))
    dec.load_state_dict({k: torch.from_numpy(v) for k, v in golden_params(g).items()})
    f, c, t = torch.from_numpy(g["features"]), torch.from_numpy(g["captions"]), torch.from_numpy(g["targets"])
    dec.zero_grad()
    loss = torch.nn.CrossEntropyLoss()(dec(f, c, g["lengths"].tolist()), t)
    loss.backward()
    assert abs(float(loss) - g["f32.loss"]) / g["f32.loss"] < 1e-6
    for k, p in dec.named_parameters():
        assert rel(p.grad.numpy(), g[f"f32.grad.{k}"]) < 1e-5, k
    np.testing.assert_array_equal(RA.sample_keepdim(dec.eval(), f).numpy(), g["f32.greedy_ids"])


def test_reference_arm_file_is_unmodified():
    import hashlib
    import os
    from oracle import ref_arm as RA
    if not RA.available() and RA.build_ref() is None:
        pytest.skip("oracle/_ref absent")
    want = open(RA.REF_SHA).read().split()[0]
    assert hashlib.sha256(open(RA.REF_FILE, "rb").read()).hexdigest() == want
    if os.path.isfile(RA.REF_SRC):
        assert open(RA.REF_SRC, "rb").read() == open(RA.REF_FILE, "rb").read()
