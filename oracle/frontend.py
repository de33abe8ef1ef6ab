"""TEST INFRASTRUCTURE (oracle): CPU restatement of the clip front-end of /root/reference/inference.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.

normalize_audio : inference.py:234
prepare_style_clip : inference.py:139-184 (query_for_motion_coeff without the pickle loading)
Pinned against the reference's own dependencies: numpy mean/std and scipy.interpolate.interp1d, which is
literally what inference.py:234 and :165-169 call (tests/test_frontend.py).
"""
import numpy as np


def normalize_audio(audio):
    audio = np.asarray(audio, dtype=np.float32)
    return (audio - audio.mean()) / (audio.std() + 1e-5)


def resample_linear(x, rows_out):
    """interp1d(kind='linear', axis=0) between the two unit-interval grids, written out."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    pos = np.linspace(0, 1, rows_out) * (n - 1)
    lo = np.clip(np.floor(pos).astype(np.int64), 0, max(n - 2, 0))
    frac = (pos - lo)[:, None]
    hi = np.minimum(lo + 1, n - 1)
    return x[lo] + frac * (x[hi] - x[lo])


def prepare_style_clip(expression_coef, head_rot, stats, original_fps=30, target_fps=25):
    exp = (np.asarray(expression_coef) - np.asarray(stats['exp_mean'])) / (np.asarray(stats['exp_std']) + 1e-9)
    rot = (np.asarray(head_rot) - np.asarray(stats['pose_mean'])) / (np.asarray(stats['pose_std']) + 1e-9)
    if original_fps is not None and original_fps != target_fps:
        new = int(round(exp.shape[0] / original_fps * target_fps))
        exp, rot = resample_linear(exp, new), resample_linear(rot, new)
    return np.concatenate([exp, rot], axis=1)[None].astype(np.float32)
