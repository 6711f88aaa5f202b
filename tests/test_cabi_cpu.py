"""CPU-side checks (no GPU): the C-ABI library builds/loads and exports exactly what include/snt_b200.h
declares; host-side logic of the boundary; and that the product path refuses to run without CUDA
(there is no CPU fallback to fall into)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def snt():
    import __graft_entry__ as ge
    ge.build()
    import show_and_tell_b200 as snt
    return snt


def _declared():
    src = open(os.path.join(ROOT, "include", "snt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(snt):
    lib = ctypes.CDLL(snt._lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/snt_b200.h but not exported"
    assert sorted(snt._lib.SIGNATURES) == names       # the ctypes table binds every declared entry point
    out = subprocess.run(["nm", "-D", "--defined-only", snt._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T snt_" in l)
    assert exported == names


def test_abi_version_and_error_string(snt):
    l = snt._lib.lib()
    assert l.snt_abi_version() == 1
    # argument validation happens before any CUDA call: bad precision -> SNT_EINVAL with a message
    rc = l.snt_lstm_fwd(7, None, 8, 8, None, None, None, None, None, 1, None, None, None, None, None, 0, None)
    assert rc == -1 and b"prec" in l.snt_last_error()
    assert l.snt_lstm_workspace_bytes(0, 100, 10, 8, 16) > 0
    assert l.snt_lstm_workspace_bytes(5, 100, 10, 8, 16) == -1
    assert l.snt_caption_trim(None, 4, 20, 2, 0, None, None, None) == -1 and b"NULL" in l.snt_last_error()
    assert l.snt_caption_trim(None, 0, 20, 2, 0, None, None, None) == 0      # empty batch: nothing to launch
    bad = (ctypes.c_int32 * 2)(1, 2)
    rc = l.snt_embed_bwd_plan(None, 4, ctypes.cast(bad, ctypes.c_void_p), 2, 10, None, 0, None)
    assert rc == -1 and b"non-increasing" in l.snt_last_error()
    ok = (ctypes.c_int32 * 2)(2, 1)
    rc = l.snt_embed_pack_bwd_planned(None, None, 4, ctypes.cast(ok, ctypes.c_void_p), 2, 2, 8, 10, None, None, None, 0,
                                      None)
    assert rc == -1 and b"bad arguments" in l.snt_last_error()          # dx is NULL
    bs = (ctypes.c_int32 * 3)(2, 3, 1)   # not non-increasing
    rc = l.snt_embed_pack_fwd(None, None, None, 4, ctypes.cast(bs, ctypes.c_void_p), 3, 8, 10, None, None, None)
    assert rc == -1 and b"non-increasing" in l.snt_last_error()


def test_batch_sizes_from_lengths(snt):
    f = snt.ops.batch_sizes_from_lengths
    assert f([3, 3, 1]).tolist() == [3, 2, 2]
    assert f([1]).tolist() == [1]
    assert f([20, 1]).tolist() == [2] + [1] * 19
    for bad in ([2, 3], [2, 0], []):
        with pytest.raises(RuntimeError):
            f(bad)
    with pytest.raises(RuntimeError):
        f([5, 2], max_steps=4)
    from oracle import snt_oracle as O
    rng = np.random.default_rng(0)
    l = snt.synthetic.make_lengths(257, rng)
    T, bs, off = O.pack_info(l)
    assert f(l).tolist() == bs.tolist()


def test_batch_sizes_cache_is_transparent(snt):
    """Repeated `lengths` lists hit a small host-side cache: same values, validation still raises, callers cannot
    corrupt the cached array, and the SM-reserve knob of the C ABI round-trips without a GPU."""
    f = snt.ops.batch_sizes_from_lengths
    a = f([5, 3, 3, 1])
    b = f([5, 3, 3, 1])
    assert a is b and a.tolist() == [4, 3, 3, 1, 1] and not a.flags.writeable
    assert f((5, 3, 3, 1)).tolist() == a.tolist()
    assert f([5, 3, 3, 1], max_steps=7).tolist() == a.tolist()
    with pytest.raises(RuntimeError):
        f([5, 3, 3, 1], max_steps=4)
    with pytest.raises(RuntimeError):
        f([3, 5])
    L = snt._lib.lib()
    assert L.snt_set_sm_reserve(12) == 0 and L.snt_set_sm_reserve(0) == 12


def test_no_cpu_fallback(snt):
    dec = snt.DecoderRNN(8, 16, 23, 1)
    feats, caps = torch.randn(2, 8), torch.ones(2, 3, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec(feats, caps, [3, 2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dec.sample(feats)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        snt.EncoderCNN(8, backbone=False).forward_pooled(torch.randn(4, 2048))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "show-and-tell_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "snt_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, fn


def test_state_dict_matches_reference_layout(snt):
    dec = snt.DecoderRNN(8, 16, 23, 2)
    sd = dec.state_dict()
    assert sd["embed.weight"].shape == (23, 8)
    assert sd["lstm.weight_ih_l0"].shape == (64, 8) and sd["lstm.weight_ih_l1"].shape == (64, 16)
    assert sd["lstm.weight_hh_l1"].shape == (64, 16) and sd["lstm.bias_hh_l0"].shape == (64,)
    assert sd["linear.weight"].shape == (23, 16) and sd["linear.bias"].shape == (23,)
    assert float(sd["linear.bias"].abs().max()) == 0.0 and float(sd["embed.weight"].abs().max()) <= 0.1
    enc = snt.EncoderCNN(8, backbone=False)
    keys = set(enc.state_dict())
    assert {"resnet.fc.weight", "resnet.fc.bias", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var",
            "bn.num_batches_tracked"} <= keys
    assert enc.bn.momentum == 0.01
    g = np.load(os.path.join(ROOT, "tests", "golden", "dec_l2_b.npz"))
    ref_keys = sorted(k[len("param."):] for k in g.files if k.startswith("param."))
    dec2 = snt.DecoderRNN(int(g["E"]), int(g["H"]), int(g["V"]), int(g["L"]))
    assert sorted(dec2.state_dict()) == ref_keys      # the reference module's own state_dict keys
    for k in ref_keys:
        assert tuple(dec2.state_dict()[k].shape) == g["param." + k].shape


def test_host_call_sequence_with_stub_binding(snt, monkeypatch):
    """The Python host side (ops.py autograd glue) enqueues the stage entry points in the documented order.  The
    ctypes binding is replaced by a recorder, so no kernel runs and the tensor contents are meaningless: this checks
    call order and that every parameter receives a gradient tensor, not numerics (those are the -m gpu tests)."""
    from show_and_tell_b200 import ops
    calls = []

    class Sizes:
        def __getattr__(self, name):
            if name.endswith("workspace_bytes"):
                return lambda *a: 4096
            raise AttributeError(name)

    monkeypatch.setattr(ops._lib, "lib", lambda: Sizes())
    monkeypatch.setattr(ops, "call", lambda name, *a: calls.append(name))
    monkeypatch.setattr(ops, "require_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "workspace", lambda nb, dev: torch.empty(max(int(nb), 1), dtype=torch.uint8))
    monkeypatch.setattr(ops, "stream_ptr", lambda: None)
    torch.manual_seed(0)
    enc, dec = snt.EncoderCNN(16, backbone=False), snt.DecoderRNN(16, 24, 50, 2)
    b = snt.synthetic.make_batch(6, 50, seed=1, pooled_dim=2048)
    pooled, caps = torch.from_numpy(b["pooled"]), torch.from_numpy(b["captions"])
    tg = torch.from_numpy(snt.synthetic.pack_host(b["captions"], b["lengths"]))
    dec.loss(enc.forward_pooled(pooled), caps, b["lengths"], tg).backward()
    assert calls == ["snt_head_fwd", "snt_embed_pack_fwd", "snt_lstm_fwd", "snt_lstm_fwd", "snt_vocab_ce_fwd",
                     "snt_vocab_ce_bwd", "snt_lstm_bwd", "snt_lstm_bwd", "snt_embed_pack_bwd", "snt_head_bwd"]
    params = [p for m in (enc, dec) for p in m.parameters() if p.requires_grad]
    assert all(p.grad is not None and p.grad.shape == p.shape for p in params)
    calls.clear()
    dec(enc.forward_pooled(pooled), caps, b["lengths"]).sum().backward()          # strict drop-in: logits materialised
    assert calls[4:6] == ["snt_linear_fwd", "snt_linear_bwd"] and calls[-1] == "snt_head_bwd" and len(calls) == 10
    calls.clear()
    dec.eval().sample(torch.randn(3, 16))
    ids = torch.zeros(3, 20, dtype=torch.int64)
    ops.trim_captions(ids)
    assert calls == ["snt_greedy_decode", "snt_caption_trim"]
    # the data-parallel step object (world size 1) on top: forward, backward, then ONE fused clip + Adam launch
    from show_and_tell_b200 import parallel
    calls.clear()
    st = parallel.DataParallelStep(enc.train(), dec.train())
    st.step(pooled, caps, b["lengths"], tg)
    assert calls[0] == "snt_head_fwd" and calls[-2:] == ["snt_head_bwd", "snt_clamp_adam_multi"] and st.t == 1
    assert len(st.m) == len(params) and all(m.shape == p.shape for m, p in zip(st.m, params))


def test_experimental_switches_host_logic_with_stub_streams(snt, monkeypatch):
    """SNT_TAIL_OVERLAP / SNT_EMB_PLAN_EARLY (off by default, first GPU run pending): their Python control flow through
    fake streams and a recording binding - the plan call precedes the loss backward, the planned variant replaces the
    one-call embedding backward, the head backward is fenced by two events."""
    import contextlib
    from show_and_tell_b200 import ops
    calls, log = [], []

    class Sizes:
        def __getattr__(self, name):
            if name.endswith("workspace_bytes"):
                return lambda *a: 4096
            raise AttributeError(name)

    class FakeEvent:
        def record(self, stream=None):
            log.append("record")

    class FakeStream:
        def __init__(self, *a, **k):
            pass

        def wait_event(self, ev):
            log.append("wait_event")

        def wait_stream(self, st):
            log.append("wait_stream")

    monkeypatch.setattr(ops._lib, "lib", lambda: Sizes())
    monkeypatch.setattr(ops, "call", lambda name, *a: calls.append(name))
    monkeypatch.setattr(ops, "require_cuda", lambda *t: None)
    monkeypatch.setattr(ops, "workspace", lambda nb, dev: torch.empty(max(int(nb), 1), dtype=torch.uint8))
    monkeypatch.setattr(ops, "stream_ptr", lambda: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    monkeypatch.setattr(ops, "TAIL_OVERLAP", True)
    monkeypatch.setattr(ops, "EMB_PLAN_EARLY", True)
    monkeypatch.setattr(ops, "_side_streams", {})
    torch.manual_seed(0)
    enc, dec = snt.EncoderCNN(16, backbone=False), snt.DecoderRNN(16, 24, 50, 1)
    b = snt.synthetic.make_batch(6, 50, seed=1, pooled_dim=2048)
    pooled, caps = torch.from_numpy(b["pooled"]), torch.from_numpy(b["captions"])
    tg = torch.from_numpy(snt.synthetic.pack_host(b["captions"], b["lengths"]))
    dec.loss(enc.forward_pooled(pooled), caps, b["lengths"], tg).backward()
    assert calls == ["snt_head_fwd", "snt_embed_pack_fwd", "snt_lstm_fwd", "snt_vocab_ce_fwd", "snt_embed_bwd_plan",
                     "snt_vocab_ce_bwd", "snt_lstm_bwd", "snt_embed_pack_bwd_planned", "snt_head_bwd"]
    # plan: wait_stream + record; dx complete: record; planned call: wait_event; head: wait_event, record, wait_event
    assert log == ["wait_stream", "record", "record", "wait_event", "wait_event", "record", "wait_event"]
    assert all(p.grad is not None for m in (enc, dec) for p in m.parameters() if p.requires_grad)
    assert ops._tail is None
