"""Multi-GPU check as a pytest (needs >= 2 GPUs; skipped on a 1-GPU test box, where tests/test_gpu_round2.py::
test_row_reductions_do_not_depend_on_the_gpu_count checks the same geometry shard by shard on one GPU): the NVLink peer-memory
exchange against the ncclAllGather fallback and against a single-GPU run, bit for bit (tools/p2p_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_memory_allreduce_matches_nccl_and_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tools", "p2p_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "p2p_check ok" in r.stdout
