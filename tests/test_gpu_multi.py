"""Multi-GPU parity (needs >= 2 B200s on one box: `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi`).
One process per GPU (torchrun), LP ranges sharded, remote messages as NVLink peer stores from inside k_step into
per-sender staging rings of the receiver (routed by k_deliver), GVT by the peer-memory mailbox round of its last block.  The gathered per-LP committed counts / trace hashes must equal the CPU oracle's,
i.e. the committed trace is identical for every number of GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "multi_rank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-3000:]); sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert r.stdout.count(": OK") >= 5 and "MISMATCH" not in r.stdout
