"""Clip-sharded multi-GPU driver: one process per GPU, full weight replica, NO collective on the data
path (clips never interact: inference.py:34-75 handles one clip; windows of a clip are sequential).

Clips are assigned in contiguous blocks by global clip id; every per-clip random input (x_T, step noise,
style eps) is keyed by the GLOBAL clip id, so a clip's result does not depend on the number of GPUs.
The only communication is the host-side gather of the finished codes / vertices to rank 0
(`torch.distributed.gather_object` over whatever backend is initialised: NCCL world on the GPU box,
gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world: int):
    """Contiguous block [lo, hi) of ceil(n/world) clips for this rank."""
    per = -(-n_clips // world)
    lo = min(rank * per, n_clips)
    return lo, min(lo + per, n_clips)


def run_sharded(n_clips, generate_fn, rank=None, world=None, gather=True):
    """generate_fn(lo, hi) -> tensor [hi-lo, ...] for global clips lo..hi-1 (computed on this rank's GPU).
    Returns the concatenated [n_clips, ...] result on rank 0 (None elsewhere) when gather=True, else the
    local block."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_clips, rank, world)
    local = generate_fn(lo, hi) if hi > lo else None
    if not gather or world == 1:
        return local
    host = None if local is None else local.detach().cpu()
    parts = [None] * world if rank == 0 else None
    dist.gather_object((lo, host), parts, dst=0)
    if rank != 0:
        return None
    parts = sorted((p for p in parts if p[1] is not None), key=lambda p: p[0])
    return torch.cat([p[1] for p in parts], 0)
