"""One process per GPU: placement from the torchrun environment and the NCCL unique-id exchange.

torch.distributed is plumbing only: rank 0 creates the NCCL unique id through the C ABI and broadcasts the 128
bytes to the other ranks; the per-iteration elite all-gather itself is issued by libicem_b200 on the planner's
stream (csrc/comm.cuh)."""
import os


def default_placement():
    """(device, world_size, rank) from RANK / LOCAL_RANK / WORLD_SIZE (torchrun); single GPU otherwise."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    return dev, ws, rank


def broadcast_bytes(payload: bytes, src=0) -> bytes:
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun / call init_process_group)")
    obj = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


def init_planner_comm(planner):
    import torch.distributed as dist
    uid = planner.comm_unique_id() if dist.get_rank() == 0 else b""
    uid = broadcast_bytes(uid, src=0)
    planner.comm_init(uid)


def shard_bounds(n_global: int, world_size: int, rank: int):
    """Contiguous ownership of global trajectory indices: rank r owns
    [r*ceil(N/R), min(N, (r+1)*ceil(N/R)))."""
    chunk = -(-n_global // world_size)
    lo = min(n_global, rank * chunk)
    hi = min(n_global, lo + chunk)
    return lo, hi


def agree_on_seed(seed):
    """All ranks of a sharded plan must draw the same Philox stream (the population is split by GLOBAL trajectory
    index): rank 0's seed wins."""
    import struct
    payload = broadcast_bytes(struct.pack("<q", int(seed)), src=0)
    return struct.unpack("<q", payload)[0]


def assert_same_on_all_ranks(array, what="start state"):
    """The ranks refit on elites gathered from all of them, which is only meaningful when every rank rolled out
    from the same start state: compare a checksum across ranks (one tiny all-gather)."""
    import hashlib

    import torch.distributed as dist
    import numpy as np
    digest = hashlib.sha256(np.ascontiguousarray(array, dtype=np.float64).tobytes()).hexdigest()
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, digest)
    if len(set(out)) != 1:
        raise RuntimeError(f"sharded planner: the {what} differs between ranks (rank 0 {out[0][:12]}..., "
                           f"rank {dist.get_rank()} {digest[:12]}...)")
