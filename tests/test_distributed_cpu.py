"""CPU tests of the multi-rank host logic (world_size 2, gloo): placement from the torchrun environment, the
unique-id broadcast used to bootstrap the library's NCCL communicator, and the sharding arithmetic -- plus the
algebraic property the device merge relies on: the global top-k equals the top-k of the per-shard top-k records."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from icem_b200 import distributed as D
dist.init_process_group("gloo")
dev, ws, rank = D.default_placement()
assert ws == 2 and rank == dist.get_rank() and dev == int(os.environ["LOCAL_RANK"])
payload = bytes(range(128)) if rank == 0 else b""
got = D.broadcast_bytes(payload, src=0)
assert got == bytes(range(128)), (rank, got[:4])
# shards tile the population for every CEM iteration
n = 262144
for it in range(3):
    lo, hi = D.shard_bounds(n, ws, rank)
    sizes = [None, None]
    dist.all_gather_object(sizes, (lo, hi))
    assert sizes[0][0] == 0 and sizes[0][1] == sizes[1][0] and sizes[1][1] == n, sizes
    n = max(20, int(n / 1.25))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_bootstrap_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2


def test_shard_bounds_match_library_partition():
    """icem_b200.distributed.shard_bounds restates csrc/planner.cu::build_plan (chunk = ceil(N_i / R))."""
    from icem_b200.distributed import shard_bounds
    for n in (2, 3, 40, 4096, 16384, 262144, 209715, 167772, 7):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = shard_bounds(n, world, r)
                assert 0 <= lo <= hi <= n
                covered += list(range(lo, hi)) if n < 100 else [lo, hi]
            if n < 100:
                assert covered == list(range(n))
            else:
                assert covered[0] == 0 and covered[-1] == n
                assert all(covered[2 * i + 1] == covered[2 * i + 2] for i in range(world - 1))


def test_topk_of_shard_topk_is_global_topk():
    rs = np.random.RandomState(0)
    from icem_b200.distributed import shard_bounds
    for n, world, k in ((1000, 8, 10), (37, 4, 10), (262147, 8, 10), (64, 2, 32)):
        c = np.round(rs.randn(n) * 3, 1).astype(np.float32)      # many ties: (cost, index) order must break them
        ref = np.argsort(c, kind="stable")[:k]
        recs = []
        for r in range(world):
            lo, hi = shard_bounds(n, world, r)
            loc = np.argsort(c[lo:hi], kind="stable")[:k] + lo
            recs += list(loc)
        recs = np.array(recs)
        merged = recs[np.lexsort((recs, c[recs]))][:k]
        np.testing.assert_array_equal(merged, ref)
