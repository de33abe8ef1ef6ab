"""Clip-sharded multi-GPU driver: one process per GPU, full weight replica, NO collective on the data
path (clips never interact: inference.py:34-75 handles one clip; windows of a clip are sequential).

Clips are assigned in contiguous blocks by global clip id; every per-clip random input (x_T, step noise,
style eps) is keyed by the GLOBAL clip id (``MSMD.sample(..., noise_seed=, clip_offset=)``), so a clip's
result does not depend on the number of GPUs.  The only communication is the gather of the finished
codes / vertices to rank 0: ``torch.distributed.gather`` of the result TENSORS (NCCL over NVLink for device
tensors on the GPU box, gloo for CPU tensors in the tests), padded to equal blocks - no pickling, so a
15 GB vertex set travels as raw bytes.
"""
import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world: int):
    """Contiguous block [lo, hi) of ceil(n/world) clips for this rank."""
    per = -(-n_clips // world)
    lo = min(rank * per, n_clips)
    return lo, min(lo + per, n_clips)


def gather_blocks(local, n_clips, rank=None, world=None, dst=0, out=None):
    """Gather the per-rank blocks (``local`` = [hi-lo, ...] of shard_range, or None for an empty shard) into one
    [n_clips, ...] tensor on rank ``dst`` (None elsewhere).  ``out`` (rank dst, optional): preallocated destination,
    e.g. pinned host memory; by default the result lives where ``local`` does."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world == 1:
        if out is not None:
            out.copy_(local)
            return out
        return local
    per = -(-n_clips // world)
    # every rank must know the trailing shape / dtype / device even if its shard is empty
    meta = [None] * world
    dist.all_gather_object(meta, None if local is None else (tuple(local.shape[1:]), str(local.dtype), str(local.device.type)))
    tail, dtype, dev = next(m for m in meta if m is not None)
    dtype = getattr(torch, dtype.split('.')[-1])
    device = local.device if local is not None else (torch.device('cuda', torch.cuda.current_device()) if dev == 'cuda' else torch.device('cpu'))
    block = torch.zeros((per,) + tail, dtype=dtype, device=device)
    if local is not None and local.shape[0]:
        block[:local.shape[0]].copy_(local)
    parts = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, parts, dst=dst)
    if rank != dst:
        return None
    if out is None:
        out = torch.empty((n_clips,) + tail, dtype=dtype, device=device)
    for r in range(world):
        lo, hi = shard_range(n_clips, r, world)
        if hi > lo:
            out[lo:hi].copy_(parts[r][:hi - lo], non_blocking=True)
    if out.device.type == 'cpu' and device.type == 'cuda':
        torch.cuda.current_stream().synchronize()
    return out


def run_sharded(n_clips, generate_fn, rank=None, world=None, gather=True, out=None):
    """generate_fn(lo, hi) -> tensor [hi-lo, ...] for global clips lo..hi-1 (computed on this rank's GPU).
    Returns the concatenated [n_clips, ...] result on rank 0 (None elsewhere) when gather=True, else the
    local block."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_clips, rank, world)
    local = generate_fn(lo, hi) if hi > lo else None
    if not gather:
        return local
    return gather_blocks(local, n_clips, rank, world, 0, out)
