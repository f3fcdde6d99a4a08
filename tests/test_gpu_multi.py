"""Multi-GPU parity on hardware: shards peer-stored into rank 0 through mr_ipc_* equal the oracle byte for byte.
Skipped automatically on a box with fewer than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_peer_store_gather_matches_oracle(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29650 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_gather_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "gather ok" in r.stdout
