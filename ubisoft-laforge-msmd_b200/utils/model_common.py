"""Drop-in for the math helpers of /root/reference/utils/model_common.py (:86-123)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class PositionalEncoding(nn.Module):
    """Sinusoid table holder (model_common.py:86-97).  The denoiser indexes ``pe`` directly
    (model.py:931); the style encoder's use of forward() adds the SINGLE row pe[:, L]
    (model_common.py:100) - that add happens inside the CUDA style encoder."""

    def __init__(self, d_model, dropout=0.1, max_len=600):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0))


def enc_dec_mask(T, S, frame_width=2, expansion=0, device='cuda'):
    """model_common.py:103-107.  True = blocked."""
    mask = torch.ones(T, S)
    for i in range(T):
        mask[i, max(0, (i - expansion) * frame_width):(i + expansion + 1) * frame_width] = 0
    return (mask == 1).to(device=device)


def pad_audio(audio, audio_unit=320, pad_threshold=80):
    """model_common.py:110-123: reflect-pad twice by side_len//2, replicate 1 if side_len is odd.
    (Host-side index plumbing; the CUDA audio encoder folds the same indexing into its loader.)"""
    batch_size, audio_len = audio.shape
    n_units = audio_len // audio_unit
    side_len = math.ceil((audio_unit * n_units + pad_threshold - audio_len) / 2)
    if side_len >= 0:
        reflect_len = side_len // 2
        if reflect_len > 0:
            audio = F.pad(audio, (reflect_len, reflect_len), mode='reflect')
            audio = F.pad(audio, (reflect_len, reflect_len), mode='reflect')
        if side_len % 2 > 0:
            audio = F.pad(audio, (1, 1), mode='replicate')
    return audio


# ---- args.json / checkpoint glue (model_common.py:9-81): host-side only, same file layout ----
def save_args(args, save_dir):
    """model_common.py:9-27: vars(args) -> <save_dir>/args.json, dropping None / 'None' entries and
    stringifying paths."""
    import json
    from pathlib import Path
    d = {k: (str(v) if isinstance(v, Path) else v) for k, v in vars(args).items() if v is not None and v != 'None'}
    with open(Path(save_dir) / 'args.json', 'w') as f:
        json.dump(d, f)


def load_args(save_dir):
    """model_common.py:52-56 / inference.py:80-84."""
    import argparse
    import json
    from pathlib import Path
    with open(Path(save_dir) / 'args.json', 'r') as f:
        return argparse.Namespace(**json.load(f))


def load_args_with_defaults(save_dir, parser):
    """model_common.py:29-50: saved values on top of the parser's current defaults."""
    import argparse
    import json
    from pathlib import Path
    with open(Path(save_dir) / 'args.json', 'r') as f:
        saved = json.load(f)
    merged = vars(parser.parse_args([]))
    merged.update(saved)
    return argparse.Namespace(**merged)


def latest_checkpoint(exp_dir):
    """model_common.py:72-77: lexicographically last checkpoints/iter_*.pt."""
    from pathlib import Path
    files = sorted((Path(exp_dir) / 'checkpoints').glob('iter_*.pt'))
    if not files:
        raise ValueError(f'No checkpoints found in {Path(exp_dir) / "checkpoints"}')
    return files[-1]
