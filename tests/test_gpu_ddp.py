"""Data-parallel training over NCCL: the gradients of a 2-rank step (bucketed all-reduce inside backward,
the non-backbone bucket overlapped with the backbone backward) equal the single-process gradients of the concatenated batch
(what DistributedDataParallel guarantees for train_spsedt.py:157-158).  Needs >= 2 GPUs (skipped on a single-GPU box; the CPU
suite covers the host-side helpers with gloo)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("graph,overlap", [(False, False), (True, False), (False, True), (True, True)])
def test_two_rank_gradients_equal_single_process(graph, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "ddp_grad_check.py")] + (["--graph"] if graph else []) + (["--overlap"] if overlap else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DDP_GRAD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
