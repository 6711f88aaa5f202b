"""2-GPU NCCL parity of the data-parallel step ("same loss, same grads after allreduce" against the
global-batch step).  Needs two visible GPUs (`gpurun --gpus 2`); skipped on a one-GPU box.  The log of the last run
under gpurun is committed as profiles/r02_multi_gpu_parity.json."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_nccl_two_rank_step_matches_global_batch_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multi_gpu_parity.py")], capture_output=True, text=True, timeout=560)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-2000:])
    rep = json.loads(lines[-1])
    assert rep["ok"] and rep["world"] == 2, rep
