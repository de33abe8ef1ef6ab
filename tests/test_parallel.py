"""CPU, world_size 2 over gloo: clip sharding + host gather give the single-process result."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_range_covers_all_clips_once():
    from msmd_b200.parallel import shard_range
    for n in (0, 1, 7, 64, 1024, 1025):
        for w in (1, 2, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            covered = [i for lo, hi in blocks for i in range(lo, hi)]
            assert covered == list(range(n))
            assert max(hi - lo for lo, hi in blocks) <= -(-n // w) if n else True


def _per_clip(lo, hi):
    """Stand-in for the GPU generation: a deterministic function of the GLOBAL clip id only, built from the
    same seeded inputs the real path uses (oracle.synth keys noise by clip id)."""
    from oracle import synth
    return torch.stack([synth.clip_xT(i)[0] * 2.0 + synth.clip_style_eps(i)[0, :67] for i in range(lo, hi)])


def _worker(rank, world, port, n_clips, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from msmd_b200.parallel import run_sharded
    res = run_sharded(n_clips, _per_clip)
    if rank == 0:
        torch.save(res, out_path)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_clips', [5, 8])
def test_two_rank_gather_matches_single_process(tmp_path, n_clips):
    out = str(tmp_path / 'gathered.pt')
    port = 29500 + (os.getpid() % 2000) + n_clips
    mp.spawn(_worker, args=(2, port, n_clips, out), nprocs=2, join=True)
    got = torch.load(out)
    assert torch.equal(got, _per_clip(0, n_clips))
