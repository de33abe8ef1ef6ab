"""Clip front-end (inference.py:139-184, :234): oracle vs the reference's own numpy / scipy calls (CPU), CUDA kernels vs
the oracle (GPU)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import frontend as F, synth


def _style_clip(frames, seed):
    g = np.random.default_rng(seed)
    stats = dict(exp_mean=g.normal(size=64).astype(np.float32), exp_std=g.uniform(0.5, 2, 64).astype(np.float32),
                 pose_mean=g.normal(size=3).astype(np.float32), pose_std=g.uniform(0.5, 2, 3).astype(np.float32))
    return g.normal(size=(frames, 64)).astype(np.float32), g.normal(size=(frames, 3)).astype(np.float32), stats


def test_oracle_resample_matches_scipy_interp1d():
    from scipy.interpolate import interp1d
    for n, m in ((120, 100), (100, 120), (2, 7), (301, 251), (50, 50), (1, 1)):
        y = np.random.default_rng(n).normal(size=(n, 5))
        if n == 1:
            assert np.allclose(F.resample_linear(y, m), np.repeat(y, m, 0))
            continue
        want = interp1d(np.linspace(0, 1, num=n), y, axis=0)(np.linspace(0, 1, num=m))        # inference.py:159-166
        assert np.abs(F.resample_linear(y, m) - want).max() < 1e-12


def test_oracle_normalize_audio():
    x = synth.clip_audio(3, 48000).numpy() * 0.3 + 0.1
    y = F.normalize_audio(x)
    assert abs(float(y.mean())) < 1e-4 and abs(float(y.std()) - 1) < 1e-3


@pytest.mark.gpu
def test_frontend_cuda_matches_oracle(built_lib):
    from msmd_b200 import inference as I
    x = torch.stack([synth.clip_audio(i, 160000) * (0.2 + i) + 0.05 * i for i in range(3)])
    got = I.normalize_audio(x.cuda())
    want = np.stack([F.normalize_audio(c.numpy()) for c in x])
    assert rel_l2(got, want) < 2e-6
    assert rel_l2(I.normalize_audio(x[1].cuda()), want[1]) < 2e-6          # 1-D clip
    for frames, fps in ((360, 30), (100, 25), (75, 15), (2, 50)):
        e, r, st = _style_clip(frames, frames)
        want = F.prepare_style_clip(e, r, st, fps, 25)
        got, shape = I.prepare_style_clip(e, r, st, 'cuda', fps, 25)
        assert got.shape == want.shape and shape.shape == (1, 100)
        assert rel_l2(got, want) < 1e-6
    with pytest.raises(Exception, match='CUDA'):
        I.normalize_audio(x)
    assert I.normalize_audio(torch.zeros(0, 16, device='cuda')).shape == (0, 16)
