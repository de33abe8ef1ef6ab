"""Oracle: audio encoder + audio feature extraction.

torch-CPU functional restatement of
  * /root/reference/utils/hubert.py:13-51 and utils/wav2vec2.py:71-119 (resampling forward),
  * the HF ``transformers`` Hubert / Wav2Vec2 *base* internals those wrappers call (pinned 4.44.2 by
    the reference's requirements.txt:124; installed 5.5.0 - same arithmetic, checked in the tests):
    feature extractor (7 bias-free Conv1d; layer 0 followed by GroupNorm(512 groups) = per-channel
    normalisation over the WHOLE clip's time axis, then GELU), feature projection (LayerNorm(512) ->
    Linear 768), positional conv (weight-normed grouped Conv1d k=128, groups=16, pad=64, drop last
    frame, GELU), encoder LayerNorm, 12 post-LN encoder layers (12 heads x 64, FFN 3072, GELU),
  * /root/reference/utils/model_common.py:110-123 (pad_audio) and model.py:250-264
    (extract_audio_feature: 2:1 linear resample + Linear 768 -> 512).
Weights come as a flat state_dict with HF key names ("audio_encoder.").
"""
import math

import torch
import torch.nn.functional as F

from .denoiser import mha  # noqa: F401  (same attention arithmetic; kept explicit below for separate q/k/v)

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)


def pad_audio(audio, audio_unit=320, pad_threshold=80):
    """model_common.py:110-123: reflect-pad TWICE by side_len//2, replicate 1 sample if side_len is odd."""
    n = audio.shape[1]
    side = math.ceil((audio_unit * (n // audio_unit) + pad_threshold - n) / 2)
    if side >= 0:
        r = side // 2
        if r > 0:
            audio = F.pad(audio, (r, r), mode='reflect')
            audio = F.pad(audio, (r, r), mode='reflect')
        if side % 2 > 0:
            audio = F.pad(audio, (1, 1), mode='replicate')
    return audio


def feature_extractor(sd, x, p='feature_extractor.'):
    """HF *FeatureEncoder with feat_extract_norm='group', conv_bias=False: [N, n] -> [N, 512, T']."""
    h = x[:, None]
    for i, (k, s) in enumerate(zip(CONV_KERNEL, CONV_STRIDE)):
        h = F.conv1d(h, sd[f'{p}conv_layers.{i}.conv.weight'], None, stride=s)
        if i == 0:
            h = F.group_norm(h, h.shape[1], sd[f'{p}conv_layers.0.layer_norm.weight'],
                             sd[f'{p}conv_layers.0.layer_norm.bias'], 1e-5)
        h = F.gelu(h)
    return h


def pos_conv_weight(sd, p='encoder.pos_conv_embed.conv.'):
    """weight_norm(dim=2): w[o,i,k] = g[k] * v[o,i,k] / ||v[:,:,k]||  (g: original0 / weight_g, v: original1 / weight_v)."""
    g = sd[p + 'parametrizations.weight.original0'] if (p + 'parametrizations.weight.original0') in sd else sd[p + 'weight_g']
    v = sd[p + 'parametrizations.weight.original1'] if (p + 'parametrizations.weight.original1') in sd else sd[p + 'weight_v']
    return g * v / v.norm(2, dim=(0, 1), keepdim=True)


def encoder(sd, h, n_layers=12, n_heads=12, p='encoder.'):
    """HF *Encoder (do_stable_layer_norm=False), eval mode: [N, T, 768] -> [N, T, 768]."""
    w = pos_conv_weight(sd)
    pos = F.conv1d(h.transpose(1, 2), w, sd[p + 'pos_conv_embed.conv.bias'], padding=64, groups=16)
    pos = F.gelu(pos[:, :, :-1]).transpose(1, 2)                          # k=128 is even: drop the last frame
    h = F.layer_norm(h + pos, (h.shape[-1],), sd[p + 'layer_norm.weight'], sd[p + 'layer_norm.bias'], 1e-5)
    d = h.shape[-1]
    dh = d // n_heads
    for l in range(n_layers):
        q_ = f'{p}layers.{l}.'
        lin = lambda name, x: F.linear(x, sd[q_ + name + '.weight'], sd[q_ + name + '.bias'])
        B, T = h.shape[:2]
        q = (lin('attention.q_proj', h) * dh ** -0.5).view(B, T, n_heads, dh).transpose(1, 2)
        k = lin('attention.k_proj', h).view(B, T, n_heads, dh).transpose(1, 2)
        v = lin('attention.v_proj', h).view(B, T, n_heads, dh).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(-1, -2), -1) @ v
        a = lin('attention.out_proj', a.transpose(1, 2).reshape(B, T, d))
        h = F.layer_norm(h + a, (d,), sd[q_ + 'layer_norm.weight'], sd[q_ + 'layer_norm.bias'], 1e-5)
        ff = lin('feed_forward.output_dense', F.gelu(lin('feed_forward.intermediate_dense', h)))
        h = F.layer_norm(h + ff, (d,), sd[q_ + 'final_layer_norm.weight'], sd[q_ + 'final_layer_norm.bias'], 1e-5)
    return h


def audio_encoder_forward(sd, input_values, output_fps=25, frame_num=None):
    """hubert.py:13-51 / wav2vec2.py:71-119: conv features -> truncate to round(frame_num*50/fps) -> linear
    interpolation to frame_num (align_corners=False) -> projection -> encoder.  Returns last_hidden_state."""
    f = feature_extractor(sd, input_values)
    if frame_num is not None:
        f = f[:, :, :round(frame_num * 50 / output_fps)]
        out_len = frame_num
    else:
        out_len = int(f.shape[2] / 50.0 * output_fps)
    f = F.interpolate(f, size=out_len, align_corners=False, mode='linear').transpose(1, 2)
    h = F.layer_norm(f, (f.shape[-1],), sd['feature_projection.layer_norm.weight'],
                     sd['feature_projection.layer_norm.bias'], 1e-5)
    h = F.linear(h, sd['feature_projection.projection.weight'], sd['feature_projection.projection.bias'])
    return encoder(sd, h)


def extract_audio_feature(sd, audio, fps, frame_num, prefix='audio_encoder.'):
    """model.py:250-264 ("Strategy 2"): encode at 2*frame_num, 2:1 linear resample, Linear 768 -> d.
    sd: MSMD.state_dict() (audio_encoder.* + audio_feature_map.*)."""
    enc = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    hs = audio_encoder_forward(enc, pad_audio(audio), fps, frame_num * 2)
    hs = F.interpolate(hs.transpose(1, 2), size=frame_num, align_corners=False, mode='linear').transpose(1, 2)
    return F.linear(hs, sd['audio_feature_map.weight'], sd['audio_feature_map.bias'])
