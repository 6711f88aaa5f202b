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
    # the step executor validates its descriptor before touching the device
    from show_and_tell_b200 import engine
    d = engine.SntStep()
    assert l.snt_step_run(ctypes.c_void_p(ctypes.addressof(d)), 15, None) == -1 and b"descriptor size" in l.snt_last_error()
    d.struct_bytes = ctypes.sizeof(engine.SntStep)
    d.prec, d.L = 1, 0
    assert l.snt_step_run(ctypes.c_void_p(ctypes.addressof(d)), 15, None) == -1 and b"L=0" in l.snt_last_error()
    d.L = 1
    assert l.snt_step_run(ctypes.c_void_p(ctypes.addressof(d)), 16, None) == -1 and b"phase" in l.snt_last_error()
    assert l.snt_step_workspace_bytes(1, 1, 64, 700, 64, 128, 1000, 2048) > 0
    assert l.snt_step_workspace_bytes(1, 1, 64, 10, 64, 128, 1000, 2048) == -1      # N < B
    assert l.snt_vocab_ce_train_workspace_bytes(0, 100, 64, 1000) == -1              # bf16 mode only
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
    # bf16 mode with a backward to follow: the stored-numerator loss (one logits contraction per step)
    assert calls == ["snt_head_fwd", "snt_embed_pack_fwd", "snt_lstm_fwd", "snt_lstm_fwd", "snt_vocab_ce_train_fwd",
                     "snt_vocab_ce_train_bwd", "snt_lstm_bwd", "snt_lstm_bwd", "snt_embed_pack_bwd", "snt_head_bwd"]
    calls.clear()
    dec.precision = "fp32"                                                        # fp32 mode: statistics + recompute
    dec.loss(enc.forward_pooled(pooled), caps, b["lengths"], tg).backward()
    assert calls[4:6] == ["snt_vocab_ce_fwd", "snt_vocab_ce_bwd"]
    dec.precision = "bf16"
    calls.clear()
    with torch.no_grad():                                                         # no backward: statistics only
        dec.loss(enc.forward_pooled(pooled), caps, b["lengths"], tg)
    assert calls[-1] == "snt_vocab_ce_fwd"
    for p in list(enc.parameters()) + list(dec.parameters()):
        p.grad = None
    calls.clear()
    dec.loss(enc.forward_pooled(pooled), caps, b["lengths"], tg).backward()
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


def test_step_descriptor_matches_header(snt, tmp_path):
    """engine.SntStep (ctypes) against struct snt_step of include/snt_b200.h, compiled here with gcc: same size and the
    same offsets for a field of every group."""
    import ctypes as C
    import subprocess
    from show_and_tell_b200 import engine
    src = tmp_path / "sz.c"
    fields = ["prec", "B", "batch_sizes", "bn_momentum", "w_emb", "w_ih", "b_hh", "d_w_fc", "d_w_ih", "d_features",
              "grad_scale", "loss", "ws_bytes"]
    src.write_text('#include "snt_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(void){\n'
                   'printf("%zu", sizeof(snt_step));\n' +
                   "".join(f'printf(" %zu", offsetof(snt_step, {f}));\n' for f in fields) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got[0] == C.sizeof(engine.SntStep)
    assert got[1:] == [getattr(engine.SntStep, f).offset for f in fields]


def test_engine_batch_sizes_and_flat_params(snt):
    """Host logic of the native step: batch_sizes (bincount form) == the broadcast form of ops, pack errors as torch
    raises them; FlatParams lays tensors out in readiness order, keeps module values, and detects a moved module."""
    from show_and_tell_b200 import engine, ops
    rng = np.random.default_rng(0)
    for _ in range(20):
        l = np.sort(rng.integers(1, 21, size=int(rng.integers(1, 300))))[::-1]
        bs, n = engine.batch_sizes(l.tolist())
        assert n == int(l.sum()) and np.array_equal(bs, ops.batch_sizes_from_lengths(l.tolist()))
    with pytest.raises(RuntimeError, match="sorted in decreasing order"):
        engine.batch_sizes([3, 5])
    with pytest.raises(RuntimeError, match="greater than 0"):
        engine.batch_sizes([3, 0])
    with pytest.raises(RuntimeError, match="exceeds"):
        engine.batch_sizes([9, 2], max_steps=8)
    torch.manual_seed(0)
    enc, dec = snt.EncoderCNN(16, backbone=False), snt.DecoderRNN(16, 24, 50, 2)
    before = {k: v.clone() for k, v in dec.state_dict().items()}
    flat = engine.FlatParams(enc, dec)
    assert flat.names[:2] == ["linear.weight", "linear.bias"] and flat.names[2].endswith("_l1")
    assert flat.names[-5] == "embed.weight" and flat.names[-1] == "encoder.bn.bias"
    lo, hi = flat.bucket_range["early"]
    assert lo == 0 and hi == flat.bucket_range["mid"][0] and flat.bucket_range["late"][1] == flat.numel
    assert all(o % 64 == 0 for o in flat.offsets) and flat.intact()
    for k, v in dec.state_dict().items():
        assert torch.equal(v, before[k])
    dec.linear.weight.data.add_(1.0)                                   # the module writes through to the flat buffer
    assert float(flat.p[0]) == float(before["linear.weight"].reshape(-1)[0] + 1.0)
    dec.load_state_dict(before)                                        # in-place copy: still on the flat buffer
    assert flat.intact() and float(flat.p[0]) == float(before["linear.weight"].reshape(-1)[0])
    flat.attach_grads()
    flat.g.fill_(2.0)
    assert float(dec.embed.weight.grad.sum()) == 2.0 * dec.embed.weight.numel()
    dec.double()                                                       # moved off the buffer
    assert not flat.intact()
    with pytest.raises(snt._lib.SntError, match="no CPU fallback"):
        engine.StepEngine(engine.FlatParams(None, snt.DecoderRNN(16, 24, 50, 1)))
