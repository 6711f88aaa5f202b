"""The native step executor (snt_step_run through engine.StepEngine / parallel.DataParallelStep) against the autograd
path of ops.py on the same modules (bit for bit where the two issue the same kernels), against torch's own Adam, on
ragged batches that change every step, and on batches wider than one co-resident group of the persistent recurrence
kernels (B > 1024: BASELINE.json configs[4] puts 2048 / 4096 captions on a GPU)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _pair(E, H, V, L, prec, head=True, seed=3):
    import show_and_tell_b200 as snt
    torch.manual_seed(seed)
    enc = snt.EncoderCNN(E, backbone=False, precision=prec).cuda().train() if head else None
    dec = snt.DecoderRNN(E, H, V, L, precision=prec).cuda().train()
    return enc, dec


def _autograd_step(enc, dec, inp, caps, lengths, targets):
    for m in (enc, dec):
        if m is not None:
            m.zero_grad(set_to_none=True)
    feats = enc.forward_pooled(inp) if enc is not None else inp
    loss = dec.loss(feats, caps, lengths, targets)
    loss.backward()
    out = {("encoder." + k if m is enc else k): p.grad.detach().clone()
           for m in (enc, dec) if m is not None for k, p in m.named_parameters()}
    return float(loss), out


def _native_names(name):
    return name.replace("encoder.resnet.fc.", "encoder.fc.")


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
@pytest.mark.parametrize("B,E,H,V,L,head", [(64, 64, 128, 1000, 1, True), (300, 128, 512, 1200, 2, True),
                                            (37, 64, 256, 777, 1, False), (1, 64, 128, 500, 1, False)])
def test_native_step_equals_autograd_path(prec, B, E, H, V, L, head):
    """Same kernels in the same order on the same data: loss and every gradient are bit-identical (the head's running
    statistics too); targets given or gathered on the device make no difference."""
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    if head and B == 1:
        pytest.skip("BatchNorm needs more than one row")
    b = snt.synthetic.make_batch(B, V, embed=E, seed=5, pooled_dim=2048)
    inp = _t(b["pooled"] if head else b["features"])
    caps, tg = _t(b["captions"]), _t(snt.synthetic.pack_host(b["captions"], b["lengths"]))
    enc, dec = _pair(E, H, V, L, prec, head)
    loss_a, g_a = _autograd_step(enc, dec, inp, caps, b["lengths"], tg)
    stats_a = None if enc is None else (enc.bn.running_mean.clone(), enc.bn.running_var.clone())
    enc2, dec2 = _pair(E, H, V, L, prec, head)
    st = parallel.DataParallelStep(enc2, dec2, optimizer=False)
    for targets in (tg, None):
        if enc2 is not None:
            enc2.bn.reset_running_stats()
        loss_n = float(st.step(inp, caps, b["lengths"], targets))
        assert loss_n == loss_a
        for k, v in g_a.items():
            g = st.flat.grad(_native_names(k))
            if k == "linear.weight" and st.engine.overlaps_dw_out():
                # the executor runs this contraction beside the BPTT from a narrower grid: another K split, same products
                assert rel_err(g.cpu().numpy(), v.cpu().numpy()) < 1e-4, k
            else:
                assert torch.equal(g, v), k
        if enc2 is not None:
            assert torch.equal(enc2.bn.running_mean, stats_a[0]) and torch.equal(enc2.bn.running_var, stats_a[1])
    assert dec2.linear.weight.grad is st.flat.grad("linear.weight")      # param.grad is the flat buffer's slice


def test_native_step_dfeatures_without_head():
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import engine
    B, E, H, V = 50, 64, 128, 600
    b = snt.synthetic.make_batch(B, V, embed=E, seed=8)
    feats, caps = _t(b["features"]), _t(b["captions"])
    _, dec = _pair(E, H, V, 1, "fp32", head=False)
    f = feats.clone().requires_grad_(True)
    dec.loss(f, caps, b["lengths"], _t(snt.synthetic.pack_host(b["captions"], b["lengths"]))).backward()
    eng = engine.StepEngine(engine.FlatParams(None, dec))
    dfe = torch.empty(B, E, device="cuda")
    eng.prepare(feats, caps, b["lengths"])
    eng.d.d_features = dfe.data_ptr()
    eng.run()
    assert torch.equal(dfe, f.grad)


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16", 2e-5)])
def test_native_adam_matches_torch_adam_on_ragged_batches(prec, tol):
    """Six steps over six DIFFERENT ragged batches (new lengths, new caption width every step): the flat-buffer
    clip + Adam of the stepper against torch.optim.Adam fed with the SAME gradients (the flat buffer's slices)."""
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    E, H, V, L = 64, 128, 900, 2
    enc, dec = _pair(E, H, V, L, prec)
    st = parallel.DataParallelStep(enc, dec, lr=2e-3, grad_clip=0.1)
    shadow = [p.detach().clone().requires_grad_(True) for p in st.params]
    opt = torch.optim.Adam(shadow, lr=2e-3)
    rng = np.random.default_rng(0)
    for it in range(6):
        B = int(rng.integers(20, 200))
        b = snt.synthetic.make_batch(B, V, seed=20 + it, pooled_dim=2048,
                                     lengths=np.sort(rng.integers(1, 8 + 3 * it, size=B))[::-1])
        loss = st.step(_t(b["pooled"]), _t(b["captions"]), b["lengths"])
        assert np.isfinite(float(loss))
        for sp, g in zip(shadow, st.flat.gviews):
            sp.grad = g.clone().clamp_(-0.1, 0.1)
        opt.step()
        for sp, p, n in zip(shadow, st.params, st.flat.names):
            assert rel_err(p.detach().cpu().numpy(), sp.detach().cpu().numpy()) < tol, (it, n)
    assert st.t == 6 and int(enc.bn.num_batches_tracked) == 6


@pytest.mark.parametrize("B", [1100, 2048])
def test_wide_batch_runs_the_persistent_recurrence_in_groups(B):
    """B > 1024: the persistent kernels loop over groups of co-resident row blocks.  Checked against the per-step
    kernels (SNT_NO_PERSISTENT=1) on the same data: the two paths share the arithmetic up to the gate functions'
    approximation, well inside the bf16 tolerance; and against themselves for determinism."""
    import os
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    E, H, V = 64, 256, 1000
    b = snt.synthetic.make_batch(B, V, embed=E, seed=9)
    feats, caps = _t(b["features"]), _t(b["captions"])

    def run():
        _, dec = _pair(E, H, V, 1, "bf16", head=False, seed=4)
        st = parallel.DataParallelStep(None, dec, optimizer=False)
        loss = float(st.step(feats, caps, b["lengths"]))
        return loss, {n: st.flat.grad(n).clone() for n in st.flat.names}

    l1, g1 = run()
    l2, g2 = run()
    assert l1 == l2 and all(torch.equal(g1[k], g2[k]) for k in g1)
    os.environ["SNT_NO_PERSISTENT"] = "1"
    try:
        l3, g3 = run()
    finally:
        os.environ.pop("SNT_NO_PERSISTENT")
    assert abs(l1 - l3) / l3 < 1e-3
    for k in g1:
        assert rel_err(g1[k].cpu().numpy(), g3[k].cpu().numpy()) < 1e-2, k


def test_stage_profile_reports_every_stage():
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    enc, dec = _pair(64, 128, 1000, 1, "bf16")
    st = parallel.DataParallelStep(enc, dec)
    b = snt.synthetic.make_batch(128, 1000, seed=2, pooled_dim=2048)
    st.engine.profile(True)
    st.step(_t(b["pooled"]), _t(b["captions"]), b["lengths"])
    prof = st.engine.profile_read()
    st.engine.profile(False)
    assert all(prof[k] > 0 for k in ("head_fwd", "embed_pack_fwd", "lstm_fwd", "vocab_ce_fwd", "vocab_ce_bwd", "lstm_bwd",
                                     "embed_pack_bwd", "head_bwd")), prof


def test_stepper_refuses_modules_moved_after_construction():
    import show_and_tell_b200 as snt
    from show_and_tell_b200 import parallel
    enc, dec = _pair(64, 128, 500, 1, "bf16")
    st = parallel.DataParallelStep(enc, dec)
    dec.float().cpu().cuda()
    b = snt.synthetic.make_batch(8, 500, seed=2, pooled_dim=2048)
    with pytest.raises(RuntimeError, match="flat buffer"):
        st.step(_t(b["pooled"]), _t(b["captions"]), b["lengths"])
