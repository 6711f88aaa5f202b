"""ctypes binding of the C-ABI library (include/snt_b200.h).  No torch types cross this boundary: raw device
pointers, sizes, a host int32 batch_sizes array and the CUDA stream handle.

There is NO fallback: if `libsnt_b200.so` is missing (or a call fails) this raises.  Build it with
`python show-and-tell_b200/build.py` or `__graft_entry__.build()`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsnt_b200.so")

PREC_FP32, PREC_BF16 = 0, 1
PREC = {"fp32": PREC_FP32, "bf16": PREC_BF16}

_vp, _i64, _i32, _f32, _f64, _int = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_double, C.c_int
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/snt_b200.h one to one
SIGNATURES = {
    "snt_abi_version": (_int, []),
    "snt_last_error": (C.c_char_p, []),
    "snt_device_query": (_int, [_int, C.POINTER(_int), C.POINTER(_int), C.POINTER(_i64)]),
    "snt_read_flags": (_int, [C.POINTER(_int), _int, _vp]),
    "snt_launch_count": (_i64, [_int]),
    "snt_set_sm_reserve": (_int, [_int]),
    "snt_gemm_f32": (_int, [_int, _int, _i64, _i64, _i64, _f32, _vp, _i64, _vp, _i64, _f32, _vp, _i64, _vp, _vp]),
    "snt_gemm_bf16": (_int, [_int, _int, _i64, _i64, _i64, _f32, _vp, _i64, _vp, _i64, _f32, _vp, _i64, _int, _vp, _vp]),
    "snt_cast_bf16": (_int, [_vp, _vp, _i64, _vp]),
    "snt_head_workspace_bytes": (_i64, [_int, _i64, _i64, _i64]),
    "snt_head_fwd": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _f32, _f32, _i64, _i64, _i64,
                            _vp, _vp, _vp, _vp, _i64, _vp]),
    "snt_head_bwd": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp,
                            _vp, _i64, _vp]),
    "snt_embed_pack_fwd": (_int, [_vp, _vp, _vp, _i64, _vp, _int, _i64, _i64, _vp, _vp, _vp]),
    "snt_embed_bwd_workspace_bytes": (_i64, [_i64, _i64]),
    "snt_embed_pack_bwd": (_int, [_vp, _vp, _i64, _vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp]),
    "snt_lstm_workspace_bytes": (_i64, [_int, _i64, _i64, _i64, _i64]),
    "snt_lstm_fwd": (_int, [_int, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp,
                            _vp, _i64, _vp]),
    "snt_lstm_bwd": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp,
                            _vp, _i64, _vp]),
    "snt_linear_workspace_bytes": (_i64, [_int, _i64, _i64, _i64]),
    "snt_linear_fwd": (_int, [_int, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp]),
    "snt_linear_bwd": (_int, [_int, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "snt_vocab_ce_workspace_bytes": (_i64, [_int, _i64, _i64, _i64]),
    "snt_vocab_ce_fwd": (_int, [_int, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp]),
    "snt_vocab_ce_bwd": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i64, _i64, _vp, _vp, _vp,
                                _vp, _i64, _vp]),
    "snt_vocab_ce_train_workspace_bytes": (_i64, [_int, _i64, _i64, _i64]),
    "snt_vocab_ce_train_fwd": (_int, [_int, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _i64, _vp]),
    "snt_vocab_ce_train_bwd": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i64, _i64, _vp, _vp, _vp,
                                      _vp, _i64, _vp]),
    "snt_step_workspace_bytes": (_i64, [_int, _int, _i64, _i64, _i64, _i64, _i64, _i64]),
    "snt_step_run": (_int, [_vp, _int, _vp]),
    "snt_step_overlaps_dw_out": (_int, [_int, _i64, _i64]),
    "snt_step_profile": (_int, [_int]),
    "snt_step_profile_read": (_int, [C.POINTER(_f32), _int]),
    "snt_greedy_workspace_bytes": (_i64, [_int, _i64, _i64, _i64, _i64, _int]),
    "snt_greedy_decode": (_int, [_int, _vp, _vp, _int, _pp, _pp, _pp, _pp, _vp, _vp, _vp, _vp,
                                 _i64, _i64, _i64, _i64, _int, _vp, _vp, _i64, _vp]),
    "snt_caption_trim": (_int, [_vp, _i64, _int, _i64, _i64, _vp, _vp, _vp]),
    "snt_clamp_adam": (_int, [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _f32, _f32, _i64, _vp]),
    "snt_dp_adam_shard": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _f64, _f64, _f64, _f64, _f32, _f32, _i64, _int, _vp]),
    "snt_clamp_adam_multi": (_int, [_int, _pp, _pp, _pp, _pp, C.POINTER(_i64), _f64, _f64, _f64, _f64, _f32, _f32,
                                    _i64, _vp]),
}

_lib = None
LAUNCHES = 0  # number of C-ABI compute calls made (bench.py reports kernel launches from this + ncu)


class SntError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SntError(f"{LIB_PATH} not found: build it with `python show-and-tell_b200/build.py` "
                           "(there is no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        if l.snt_abi_version() != 1:
            raise SntError("libsnt_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().snt_last_error().decode(errors="replace")
        raise SntError(f"{what} failed ({rc}): {msg}")


_profile = None  # when a list: (name, start_event, end_event) per C-ABI call, see profile_begin/profile_end


def call(name, *args):
    global LAUNCHES
    LAUNCHES += 1
    if _profile is None:
        check(getattr(lib(), name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(lib(), name)(*args), name)
    e1.record()
    _profile.append((name, e0, e1))


def profile_begin():
    """Start recording a CUDA-event pair around every C-ABI call (diagnostics; adds event overhead)."""
    global _profile
    _profile = []


def profile_end():
    """-> {entry point: (calls, total GPU milliseconds)} since profile_begin()."""
    global _profile
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _profile or []:
        n, t = out.get(name, (0, 0.0))
        out[name] = (n + 1, t + e0.elapsed_time(e1))
    _profile = None
    return out


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _raw_stream(device_index=None):
    """Raw cudaStream_t of torch's current stream.  torch.cuda.current_stream() builds a Stream object and resolves the
    device through several Python layers (~14 us per call, 16 calls per training step: measured with cProfile); the
    private accessor below is a single C call."""
    if device_index is None:
        device_index = torch._C._cuda_getDevice()
    return torch._C._cuda_getCurrentRawStream(device_index)


def stream_ptr():
    return C.c_void_p(_raw_stream())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SntError("snt_b200 ops need CUDA tensors: there is no CPU fallback")


def host_i32(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return a, a.ctypes.data_as(C.c_void_p)


_workspaces = {}


def workspace(nbytes, device):
    """A cached, growing scratch buffer per (device, stream)."""
    if nbytes < 0:
        raise SntError("workspace query failed (bad sizes or precision)")
    idx = device.index if isinstance(device, torch.device) and device.index is not None else None
    key = (idx if idx is not None else str(device), _raw_stream(idx))
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        w = None
        _workspaces.pop(key, None)
        w = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w


def read_flags(reset=True):
    f = C.c_int(0)
    check(lib().snt_read_flags(C.byref(f), 1 if reset else 0, stream_ptr()), "snt_read_flags")
    return f.value
