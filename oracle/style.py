"""Oracle: StyleEncoder_VAE2 (torch-CPU functional restatement of /root/reference/style_encoder.py:119-213).

conv(k3)+ELU+LN x2 -> add the SINGLE positional row pe[L] (PositionalEncoding.forward bug,
model_common.py:100) -> one post-LN TransformerEncoderLayer (d512, 8 heads, ff512,
GELU) -> conv+ELU+LN -> conv -> mean over time -> (mu, logvar) -> mu + eps * exp(0.5 logvar).
Dropout layers are identity in eval mode.
"""
import torch
import torch.nn.functional as F

from .denoiser import mha, sinusoid_table


def _conv_t(sd, name, x):
    """Conv1d(k=3, padding=1) over time on channels-last input [N, L, C] (style_encoder.py:137-139)."""
    return F.conv1d(x.transpose(1, 2), sd[name + '.weight'], sd[name + '.bias'], padding=1).transpose(1, 2)


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + '.weight'], sd[name + '.bias'], 1e-5)


def style_stats(sd, motion):
    """style_encoder.py:186-197 -> (mu, logvar), each [N, d_style].  sd keys as in StyleEncoder_VAE2.state_dict()."""
    L = motion.shape[1]
    x = _ln(sd, 'input_layers.5', F.elu(_conv_t(sd, 'input_layers.1', motion)))
    x = _ln(sd, 'input_layers.11', F.elu(_conv_t(sd, 'input_layers.7', x)))
    pe = sd['PE.pe'] if 'PE.pe' in sd else sinusoid_table(600, x.shape[-1]).unsqueeze(0)
    x = x + pe[:, L, :]                                                   # one row for every position
    p = 'encoder.'
    sa = mha(x, x, sd[p + 'self_attn.in_proj_weight'], sd[p + 'self_attn.in_proj_bias'],
             sd[p + 'self_attn.out_proj.weight'], sd[p + 'self_attn.out_proj.bias'], 8)
    x = _ln(sd, p + 'norm1', x + sa)
    ff = F.linear(F.gelu(F.linear(x, sd[p + 'linear1.weight'], sd[p + 'linear1.bias'])),
                  sd[p + 'linear2.weight'], sd[p + 'linear2.bias'])
    x = _ln(sd, p + 'norm2', x + ff)
    x = _ln(sd, 'output_layers.5', F.elu(_conv_t(sd, 'output_layers.1', x)))
    out = _conv_t(sd, 'output_layers.7', x).mean(dim=1)                   # style_encoder.py:191-194
    h = out.shape[1] // 2
    return out[:, :h], out[:, h:]


def style_forward(sd, motion, eps):
    """forward(do_sample=False) with the noise supplied: (mu + eps * std, mu, logvar) (style_encoder.py:196-208)."""
    mu, logvar = style_stats(sd, motion)
    return mu + eps * torch.exp(0.5 * logvar), mu, logvar
