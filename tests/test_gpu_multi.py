"""T6: N-GPU sharded plan == 1-GPU plan, bit-identical (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_plan_matches_single_gpu(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
                          os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.count("bit-identical") == 3
