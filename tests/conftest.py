import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_params(g):
    return {k[len("param."):]: g[k] for k in g if k.startswith("param.")}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def decoder_from_params(params, prec, device="cuda"):
    """A show_and_tell_b200.DecoderRNN carrying the given state_dict-keyed numpy weights."""
    import torch
    import show_and_tell_b200 as snt
    V, E = params["embed.weight"].shape
    H = params["lstm.weight_hh_l0"].shape[1]
    L = sum(1 for k in params if k.startswith("lstm.weight_ih_l"))
    dec = snt.DecoderRNN(E, H, V, L, precision=prec)
    dec.load_state_dict({k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in params.items()})
    return dec.to(device)
