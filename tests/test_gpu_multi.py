"""Two-GPU check of the sharded batch path (NCCL all-reduce of the clipped sums).  Skipped on a
one-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("margin", [None, "0"])
def test_sharded_equals_single_gpu(cuda, margin):
    """margin "0": the sharded sampler draws no redundant tiles, so position ranges that reach into a
    neighbour's slice exercise the re-draw path of the compaction kernel."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    if margin is not None:
        env["D3P_TEST_SAMPLER_MARGIN"] = margin
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "helpers", "multi_rank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MULTI_RANK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
