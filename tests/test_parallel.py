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


# ------------------------------------------------------------------------------------------------------------------
# The REAL path on GPUs: clip i sampled on a 2-GPU run (contiguous shards, gathered with run_sharded) is bit-identical
# to clip i sampled on one GPU.  Needs 2 GPUs (gpurun --gpus 2); skipped otherwise.
N_ID, T_ID, SEED_ID = 6, 12, 99


def _generate_real(lo, hi, device, noise):
    """Codes of global clips lo..hi-1: 2 windows of the real sampler (hybrid default arithmetic) + FLAME landmarks-free decode.
    noise = 'z': externally supplied per-clip z; 'philox': in-kernel noise keyed by the global clip id."""
    from helpers import make_msmd
    from msmd_b200.inference import infer_coeffs_batched
    from oracle import synth
    m, args = make_msmd(device, precision=None, n_diff_steps=T_ID)
    ids = list(range(lo, hi))
    g = lambda fn: torch.cat([fn(i) for i in ids]).to(device)
    feat = torch.stack([torch.randn(200, 512, generator=torch.Generator().manual_seed(700 + i)) for i in ids]).to(device)
    style = g(synth.clip_style_eps)
    x_T = g(synth.clip_xT)
    z = None
    if noise == 'z':
        z = [torch.stack([synth.clip_step_noise(i, w, T_ID) for i in ids], 1).to(device) for w in range(2)]
    return infer_coeffs_batched(m, args, feat, torch.zeros(len(ids), 1, 100, device=device), style, clip_len=170, cfg_scale=1.4,
                                x_T=x_T, noise=z, noise_seed=SEED_ID, clip_offset=lo)


def _gpu_worker(rank, world, port, noise, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from msmd_b200.parallel import run_sharded
    res = run_sharded(N_ID, lambda lo, hi: _generate_real(lo, hi, dev, noise))
    if rank == 0:
        torch.save(res.cpu(), out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize('noise', ['z', 'philox'])
def test_two_gpu_codes_identical_to_one_gpu(built_lib, tmp_path, noise):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    out = str(tmp_path / f'codes_{noise}.pt')
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_gpu_worker, args=(2, port, noise, out), nprocs=2, join=True)
    two = torch.load(out)
    one = _generate_real(0, N_ID, torch.device('cuda', 0), noise).cpu()
    assert two.shape == (N_ID, 170, 67) and torch.isfinite(one).all()
    assert torch.equal(two, one), float((two - one).abs().max())


@pytest.mark.gpu
def test_philox_noise_is_keyed_by_global_clip_id(built_lib):
    """One GPU: clips 2..5 sampled as their own batch (clip_offset=2) equal rows 2..5 of the 6-clip batch."""
    dev = torch.device('cuda', 0)
    whole = _generate_real(0, N_ID, dev, 'philox')
    part = _generate_real(2, 6, dev, 'philox')
    assert torch.equal(part, whole[2:6])
    other = _generate_real(0, 4, dev, 'philox')
    assert torch.equal(other, whole[:4]) and not torch.equal(whole[0], whole[1])
