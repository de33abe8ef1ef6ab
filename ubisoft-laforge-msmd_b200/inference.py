"""Drop-in for the hot-path driver of /root/reference/inference.py: ``infer_coeffs`` (inference.py:34-75).

Same signature and windowing semantics: the audio encoder sees the whole zero-padded clip once, windows
of ``n_motions`` frames are sampled sequentially, each conditioned on the previous window's last
``n_prev_motions`` frames of motion and of INPUT audio features, every window re-uses window 0's x_T
(inference.py:57-69), and the padded tail is trimmed.

``infer_coeffs_batched`` is the same loop over a batch of independent clips (the reference handles one
clip per call; clips never interact, so they are simply stacked along the batch dimension).
"""
import math

import torch
import torch.nn.functional as F


@torch.no_grad()
def infer_coeffs_batched(model, args, audio_feat, shape_coef, style_feats=None, clip_len=None, cfg_mode=None,
                         cfg_cond=None, cfg_scale=1.15, dynamic_threshold=None, x_T=None, noise=None, noise_seed=None,
                         clip_offset=0):
    """audio_feat [N, n_sub*n_motions, d] (already extracted); shape_coef [N,1,100] or [N,100];
    style_feats [N,d_style] (or a list per window); clip_len = frames to keep (<= n_sub*n_motions).
    x_T [N,n_motions,67] / noise ([T+1,N,L,67] or a list per window) are optional fixed inputs; without ``noise`` the
    step noise is the in-kernel Philox stream keyed by (noise_seed + window, global clip id = clip_offset + n), so the
    codes of a clip do not depend on which batch / GPU it was sampled in.
    Returns [N, clip_len, 67]."""
    N, total = audio_feat.shape[:2]
    L = args.n_motions
    assert total % L == 0, f'audio feature length {total} is not a multiple of n_motions={L}'
    n_sub = total // L
    clip_len = total if clip_len is None else clip_len
    n_padding_frames = total - clip_len
    coef_list = []
    prev_motion_feat = prev_audio_feat = None
    noise_T = x_T
    for i in range(n_sub):
        indicator = torch.ones((N, L), device=audio_feat.device) if args.use_indicator else None
        if indicator is not None and i == n_sub - 1 and n_padding_frames > 0:
            indicator[:, -n_padding_frames:] = 0
        audio_in = audio_feat[:, i * L:(i + 1) * L]
        style_feat = style_feats[i] if isinstance(style_feats, list) else style_feats
        z = noise[i] if isinstance(noise, list) else noise
        motion_feat, noise_T, used_audio = model.sample(
            audio_in, shape_coef, style_feat, prev_motion_feat, prev_audio_feat, noise_T, indicator=indicator,
            cfg_mode=cfg_mode, cfg_cond=cfg_cond, cfg_scale=cfg_scale, dynamic_threshold=dynamic_threshold, noise=z,
            noise_seed=None if noise_seed is None else int(noise_seed) + i, clip_offset=clip_offset)
        prev_motion_feat = motion_feat[:, -args.n_prev_motions:].clone()
        prev_audio_feat = used_audio[:, -args.n_prev_motions:]
        if i == n_sub - 1 and n_padding_frames > 0:
            motion_feat = motion_feat[:, :-n_padding_frames]
        coef_list.append(motion_feat)
    return torch.cat(coef_list, dim=1)


@torch.no_grad()
def infer_coeffs(model, args, audio, shape_coef, audio_unit, style_feats=None, n_repetitions: int = 1, cfg_mode=None,
                 cfg_cond=None, cfg_scale: float = 1.15, include_shape: bool = False, dynamic_threshold=(0, 1, 4)):
    """inference.py:34-75 (one clip, 1-D 16 kHz audio)."""
    clip_len = int(len(audio) / 16000 * args.fps)
    n_audio_samples = round(audio_unit * args.n_motions)
    n_subdivision = 1 if clip_len <= args.n_motions else math.ceil(clip_len / args.n_motions)
    n_padding_audio_samples = n_audio_samples * n_subdivision - len(audio)
    n_padding_frames = math.ceil(n_padding_audio_samples / audio_unit)
    if n_padding_audio_samples > 0:
        audio = F.pad(audio, (0, n_padding_audio_samples), value=0)
    audio_feat = model.extract_audio_feature(audio.unsqueeze(0), args.n_motions * n_subdivision)
    total = args.n_motions * n_subdivision
    rep = lambda t: t if t is None else t.expand(n_repetitions, *t.shape[1:])
    style = [rep(s) for s in style_feats] if isinstance(style_feats, list) else rep(style_feats)
    return infer_coeffs_batched(model, args, audio_feat.expand(n_repetitions, -1, -1), rep(shape_coef), style,
                                clip_len=total - max(n_padding_frames, 0), cfg_mode=cfg_mode, cfg_cond=cfg_cond,
                                cfg_scale=cfg_scale, dynamic_threshold=dynamic_threshold)


def load_model(model_root: str, model_name: str, iter_num: str, device='cuda'):
    """inference.py:85-103: <model_root>/DPT/<model_name>/{args.json, checkpoints/iter_<iter_num>.pt} ->
    (model, style_encoder, args).  The checkpoint's 'model' / 'style_enc' state_dicts load unchanged
    (same keys and shapes as the reference modules)."""
    import os
    from pathlib import Path
    from .model import get_diffusion_model
    from .style_encoder import get_style_encoder
    from .utils.model_common import load_args
    exp = Path(os.path.join(model_root, "DPT", model_name))
    model_args = load_args(exp)
    model = get_diffusion_model(model_args, device)
    ckpt = torch.load(exp / "checkpoints" / f"iter_{iter_num}.pt", map_location=device)
    style_enc = get_style_encoder(model_args, model_args.style_enc_model_style)
    style_enc.load_state_dict(ckpt['style_enc'])
    style_enc.to(device).eval()
    model.load_state_dict(ckpt['model'])
    model.eval()
    return model, style_enc, model_args


# ---------------------------------------------------------------------------------------------------------------
# Clip front-end: the numpy stages of inference.py between file I/O and the encoders,
# on the GPU through the C ABI (csrc/frontend.cu).  File decoding itself (librosa.load / pickle) stays with the caller.
def normalize_audio(audio):
    """inference.py:234: ``(audio - audio.mean()) / (audio.std() + 1e-5)`` per clip (numpy's population std).
    audio: CUDA fp32 [n] or [N, n]."""
    from . import _lib
    if audio.device.type != 'cuda':
        raise _lib.MsmdError('msmd_b200.inference.normalize_audio needs a CUDA tensor (no CPU path)')
    x = audio.detach().to(torch.float32).contiguous()
    x2 = x.reshape(1, -1) if x.ndim == 1 else x
    out = torch.empty_like(x2)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().msmd_audio_normalize(_lib.dev_ptr(x2), _lib.dev_ptr(out), x2.shape[0], x2.shape[1],
                                                   _lib.stream_ptr()))
    return out.reshape(x.shape)


def resample_linear(x, rows_out):
    """scipy ``interp1d(np.linspace(0, 1, rows_in), x, axis=0)(np.linspace(0, 1, rows_out))`` (inference.py:158-171).
    x: CUDA fp32 [rows_in, cols]."""
    from . import _lib
    if x.device.type != 'cuda':
        raise _lib.MsmdError('msmd_b200.inference.resample_linear needs a CUDA tensor (no CPU path)')
    x = x.detach().to(torch.float32).contiguous()
    if x.ndim != 2 or x.shape[0] < 1:
        raise ValueError(f'resample_linear expects [rows >= 1, cols], got {tuple(x.shape)}')
    out = torch.empty((int(rows_out), x.shape[1]), device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().msmd_resample_linear(_lib.dev_ptr(x), _lib.dev_ptr(out), x.shape[0], int(rows_out), x.shape[1],
                                                   _lib.stream_ptr()))
    return out


def prepare_style_clip(expression_coef, head_rot, coef_stats, device='cuda', original_fps=30, target_fps=25):
    """The arithmetic of ``query_for_motion_coeff`` (inference.py:108-184) on already-loaded arrays: normalise the
    expression codes and head rotations with the dataset statistics (eps 1e-9), resample original_fps -> target_fps
    by linear interpolation, concatenate.  Returns (motion_coeff [1, frames, n_exp + n_rot], shape_coef zeros [1, 100])."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device=device)
    exp = (t(expression_coef) - t(coef_stats['exp_mean'])) / (t(coef_stats['exp_std']) + 1e-9)
    rot = (t(head_rot) - t(coef_stats['pose_mean'])) / (t(coef_stats['pose_std']) + 1e-9)
    if original_fps is not None and original_fps != target_fps:
        new_frames = int(round(exp.shape[0] / original_fps * target_fps))
        exp, rot = resample_linear(exp, new_frames), resample_linear(rot, new_frames)
    return torch.cat([exp, rot], dim=1).unsqueeze(0), torch.zeros((1, 100), device=device)
