"""row-sharded decode over >= 2 GPUs: NVLink peer-store all-gather (fused into the consumer kernel) against ncclAllGather
on the same seeded weights -- identical tokens, logits within the fp32-atomics noise.  Skipped on single-GPU boxes; the
host-side shard / gather logic is covered on CPU by tests/test_shard_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_exchange_matches_nccl():
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_tp.py"), "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ALL OK") == n
