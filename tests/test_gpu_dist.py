"""Multi-GPU parity inside `pytest -m gpu` (runs when the box shows >= 2 GPUs, skipped cleanly on a 1-GPU box; the same
checks run inside every `bench.py` at N > 1 as `dist_parity`, so the driver's scaling runs carry them as well).

tests/dist_check.py under torchrun: distributed Sinkhorn (in-kernel NVLink peer exchange, the default; and the NCCL
all-reduce path) == the reference's own 2-rank fixture and == the fp64 global-batch oracle; hybrid kernel (rows beyond
shared memory) across ranks; clips sharded by rank == one process (bit for bit)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))


@pytest.mark.parametrize("path", ["p2p", "nccl"])
def test_two_rank_dist_check(path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(2, 29517 if path == "p2p" else 29518, {"TIMET_SK_P2P": "1" if path == "p2p" else "0"})
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"sinkhorn path: {path}" in r.stdout, r.stdout[-1500:]
