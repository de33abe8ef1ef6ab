"""Oracle: diffusion schedule, denoising network, CFG sampler and the windowing driver.

torch-CPU functional restatement of /root/reference/model.py (DiffusionSchedule :20-71,
MSMD.sample :282-440, DenoisingNetwork_MSMD.forward :914-996), utils/model_common.py
(PositionalEncoding :86-101, enc_dec_mask :103-107) and inference.infer_coeffs
(inference.py:34-75).  Weights come as a flat ``state_dict`` with the reference's key
names; nothing here builds nn.Modules, so the arithmetic is explicit:
post-LN decoder layers, packed q|k|v projections, exact-erf GELU, eps 1e-5.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- tables
def sinusoid_table(max_len, d_model):
    """model_common.py:90-97: pe[pos, 2i] = sin(pos * w_i), pe[pos, 2i+1] = cos(pos * w_i)."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def alignment_mask(n_prev, n_motions, width=1):
    """model.py:879-883 + model_common.py:103-107.  True = blocked.  Row 0 (person token) sees
    everything; motion row i sees memory columns [i-1-(w-1), i-1+(w-1)]."""
    T = n_prev + n_motions
    m = torch.ones(T, T)
    ex = width - 1
    for i in range(T):
        m[i, max(0, i - ex):(i + ex + 1)] = 0
    m = (m == 1)
    return torch.cat([torch.zeros(1, T, dtype=torch.bool), m], 0)      # [T+1, T]


def cosine_schedule(num_steps, s=0.008):
    """model.py:32-58 (mode='cosine'): betas clipped to [1e-4, 0.999], alpha_bar by running
    log-sum, sigmas_inflex from the posterior variance.  Returns dict of [num_steps+1] buffers."""
    x = torch.linspace(0, num_steps, num_steps + 1)
    ab = torch.cos(((x / num_steps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ab = ab / ab[0]
    betas = torch.clip(1 - (ab[1:] / ab[:-1]), 0.0001, 0.999)
    betas = torch.cat([torch.zeros(1), betas], 0)
    alphas = 1 - betas
    la = torch.log(alphas)
    for i in range(1, la.shape[0]):
        la[i] += la[i - 1]
    alpha_bars = la.exp()
    sig_flex = torch.sqrt(betas)
    sig_inflex = torch.zeros_like(sig_flex)
    for i in range(1, sig_flex.shape[0]):
        sig_inflex[i] = ((1 - alpha_bars[i - 1]) / (1 - alpha_bars[i])) * betas[i]
    sig_inflex = torch.sqrt(sig_inflex)
    return dict(betas=betas, alphas=alphas, alpha_bars=alpha_bars, sigmas_flex=sig_flex, sigmas_inflex=sig_inflex)


# ----------------------------------------------------------------------------- building blocks
def _lin(sd, name, x):
    return F.linear(x, sd[name + '.weight'], sd[name + '.bias'])


def _ln(sd, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + '.weight'], sd[name + '.bias'], eps)


def mha(x_q, x_kv, w_in, b_in, w_out, b_out, n_heads, mask=None):
    """nn.MultiheadAttention (batch_first, packed in_proj q|k|v).  mask bool [Tq,Tk], True = blocked."""
    d = x_q.shape[-1]
    dh = d // n_heads
    q = F.linear(x_q, w_in[:d], b_in[:d])
    k = F.linear(x_kv, w_in[d:2 * d], b_in[d:2 * d])
    v = F.linear(x_kv, w_in[2 * d:], b_in[2 * d:])
    B, Tq, Tk = q.shape[0], q.shape[1], k.shape[1]
    q = q.view(B, Tq, n_heads, dh).transpose(1, 2)
    k = k.view(B, Tk, n_heads, dh).transpose(1, 2)
    v = v.view(B, Tk, n_heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if mask is not None:
        s = s.masked_fill(mask, float('-inf'))
    o = torch.softmax(s, -1) @ v
    return F.linear(o.transpose(1, 2).reshape(B, Tq, d), w_out, b_out)


def decoder_layer(sd, p, x, mem, n_heads, mask):
    """nn.TransformerDecoderLayer(norm_first=False, activation='gelu'): model.py:874-878."""
    sa = mha(x, x, sd[p + 'self_attn.in_proj_weight'], sd[p + 'self_attn.in_proj_bias'],
             sd[p + 'self_attn.out_proj.weight'], sd[p + 'self_attn.out_proj.bias'], n_heads)
    x = _ln(sd, p + 'norm1', x + sa)
    ca = mha(x, mem, sd[p + 'multihead_attn.in_proj_weight'], sd[p + 'multihead_attn.in_proj_bias'],
             sd[p + 'multihead_attn.out_proj.weight'], sd[p + 'multihead_attn.out_proj.bias'], n_heads, mask)
    x = _ln(sd, p + 'norm2', x + ca)
    ff = _lin(sd, p + 'linear2', F.gelu(_lin(sd, p + 'linear1', x)))
    return _ln(sd, p + 'norm3', x + ff)


def denoiser_forward(sd, cfg, motion, audio, person, style, prev_motion, prev_audio, step, indicator=None,
                     prefix='denoising_net.', keep_separate=False):
    """model.py:914-996.  motion [N,L,67]; audio [N,L,d]; person [N,1,100+d_style]; style [N,1,d_style];
    prev_motion [N,Lp,67]; prev_audio [N,Lp,d]; step [N] long; indicator [N,L] or None.
    Returns [N, Lp+L, 67]."""
    g = lambda k: sd[prefix + k]
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    N = motion.shape[0]
    n_prev = prev_motion.shape[1]
    te = g('TE.pe')[0, step]                                                        # model.py:931
    emb = _lin(sub, 'diff_step_map.2', F.gelu(_lin(sub, 'diff_step_map.0', te))).unsqueeze(1)
    ptok = _lin(sub, 'person_proj', person) + emb                                   # :932-933
    feats = torch.cat([prev_motion, motion], 1)                                     # :940
    if cfg.use_indicator:
        ind = torch.cat([torch.zeros(N, n_prev), indicator], 1).unsqueeze(-1)       # :935-937
        feats = torch.cat([feats, ind], -1)                                         # :944
    x = torch.cat([ptok, _lin(sub, 'feature_proj', feats)], 1) + g('PE')            # :945-949
    mem = torch.cat([prev_audio, audio], 1)                                         # :955
    mask = g('alignment_mask') if (prefix + 'alignment_mask') in sd else None
    for l in range(cfg.n_layers):
        x = decoder_layer(sub, f'transformer.layers.{l}.', x, mem, cfg.n_heads, mask)
    out = _lin(sub, 'motion_dec.2', F.gelu(_lin(sub, 'motion_dec.0', x[:, 1:])))     # :961
    nb = int(cfg.num_of_basis)
    static = torch.stack([_lin(sub, f'static_feature_mapping.{b}.2',
                               F.gelu(_lin(sub, f'static_feature_mapping.{b}.0', style)))[:, 0]
                          for b in range(nb)], 1)                                   # [N, nb, 67]  :964-970
    alphas = out[:, :, -nb:]                                                         # :972
    dyn = out[:, :, :-nb]
    if keep_separate:
        return dyn, static, alphas
    face = torch.einsum('nlb,nbc->nlc', alphas, static[:, :, :-3])                   # :985-988 (use_head_alpha=False)
    pose = static[:, :, -3:].sum(1, keepdim=True).expand(-1, out.shape[1], -1)       # :989 (unweighted)
    return dyn + torch.cat([face, pose], -1)                                         # :995


# ----------------------------------------------------------------------------- sampler
def cfg_entries(sd, cfg, audio_feat, shape_feat, style_feat, cfg_mode, cfg_cond):
    """model.py:339-374: per-entry (audio, person) conditioning.  Entry 0 is the null entry."""
    N, L = audio_feat.shape[:2]
    null_style = sd['null_style_feat'].expand(N, -1, -1) if 'null_style_feat' in sd else None
    a_null = sd['null_audio_feat'].expand(N, L, -1) if 'audio' in cfg_cond else audio_feat
    if 'style' in cfg_cond:
        p_null = torch.cat([shape_feat, null_style], -1)
    else:
        p_null = torch.cat([shape_feat, style_feat], -1)
    audio_in, person_in = [a_null], [p_null]
    for cond in cfg_cond:
        if cond == 'audio':
            audio_in.append(audio_feat)
            person_in.append(p_null)
        else:
            if cfg_mode == 'independent':
                audio_in.append(a_null)
            elif cfg_mode == 'incremental':
                audio_in.append(audio_feat)
            else:
                raise NotImplementedError(f'Unknown cfg_mode {cfg_mode}')
            person_in.append(torch.cat([shape_feat, style_feat], -1))
    return audio_in, person_in


def sample(sd, cfg, audio_feat, shape_feat, style_feat, prev_motion=None, prev_audio=None, x_T=None,
           z=None, indicator=None, cfg_mode=None, cfg_cond=None, cfg_scale=1.15, flexibility=0,
           ret_traj=False, n_steps=None, denoise_fn=None, dynamic_threshold=None, separate=False, t_start=None):
    """model.py:282-440 with externally supplied noise.

    z: [T+1, N, L, 67] indexed by t (z[t] used at step t > 1; zeros at t == 1), or None -> torch.randn.
    n_steps (< T) stops early after that many steps (tests); returns the state reached.
    t_start (tests): begin at step t_start with x_T taken as x_{t_start} (teacher-forced single steps).
    Returns (x, x_T, audio_feat) like the reference; with ret_traj a dict {t: x_t}.
    dynamic_threshold = (ratio, min, max): per-sequence quantile clamp of the network output (model.py:396-402).
    separate=True follows MSMD.sample_separate (model.py:442-651, alpah_t_modification=None) and returns
    (x, x_T, audio_feat, target_dynamic, cumulative_static, alpha_traj[T_run*N, L, n_basis]).
    """
    N = audio_feat.shape[0]
    sched = {k[len('diffusion_sched.'):]: v for k, v in sd.items() if k.startswith('diffusion_sched.')}
    T = sched['betas'].shape[0] - 1
    cfg_mode = cfg_mode or cfg.cfg_mode
    if cfg_cond is None:
        cfg_cond = [c for c in cfg.guiding_conditions.split(',') if c in ('style', 'audio')]
    cfg_cond = [c for c in cfg_cond if c in ('audio', 'style')]
    if not isinstance(cfg_scale, list):
        cfg_scale = [cfg_scale] * len(cfg_cond)
    if cfg_cond:
        cfg_cond, cfg_scale = zip(*sorted(zip(cfg_cond, cfg_scale), key=lambda x: ['audio', 'style'].index(x[0])))
    if style_feat is None:
        style_feat = sd['null_style_feat'].expand(N, -1, -1)
    if shape_feat.ndim == 2:
        shape_feat = shape_feat.unsqueeze(1)
    if style_feat.ndim == 2:
        style_feat = style_feat.unsqueeze(1)
    if prev_motion is None:
        prev_motion = sd['start_motion_feat'].expand(N, -1, -1)
    if prev_audio is None:
        prev_audio = sd['start_audio_feat'].expand(N, -1, -1)
    if x_T is None:
        x_T = torch.randn(N, cfg.n_motions, 67)
    audio_in, person_in = cfg_entries(sd, cfg, audio_feat, shape_feat, style_feat, cfg_mode, cfg_cond)
    E = len(audio_in)
    audio_in = torch.cat(audio_in, 0)
    person_in = torch.cat(person_in, 0)
    pm = torch.cat([prev_motion] * E, 0)
    pa = torch.cat([prev_audio] * E, 0)
    ind = torch.cat([indicator] * E, 0) if indicator is not None else None
    st = torch.cat([style_feat] * E, 0)                        # real style for every entry (model.py:374)
    fn = denoise_fn or (lambda *a: denoiser_forward(sd, cfg, *a, keep_separate=separate))
    cum_static = torch.zeros_like(x_T)
    alpha_traj, tgt_dyn = [], None
    x = x_T
    T0 = t_start or T
    traj = {T0: x_T}
    last = T0 - n_steps if n_steps else 0
    for t in range(T0, last, -1):
        zt = (z[t] if z is not None else torch.randn_like(x)) if t > 1 else torch.zeros_like(x)
        alpha, ab, ab_prev = sched['alphas'][t], sched['alpha_bars'][t], sched['alpha_bars'][t - 1]
        sigma = sched['sigmas_flex'][t] * flexibility + sched['sigmas_inflex'][t] * (1 - flexibility)
        step = torch.full((N * E,), t, dtype=torch.long)
        res = fn(torch.cat([x] * E, 0), audio_in, person_in, st, pm, pa, step, ind)
        if separate:
            dyn, stat_b, alph = res                                     # model.py:557-575
            face = torch.einsum('nlb,nbc->nlc', alph, stat_b[:, :, :-3])
            pose = stat_b[:, :, -3:].sum(1, keepdim=True).expand(-1, dyn.shape[1], -1)
            stat = torch.cat([face, pose], -1)
            res = dyn + stat
        if dynamic_threshold:                                          # model.py:396-402 / :579-585
            q, lo, hi = dynamic_threshold
            a = res[:, -cfg.n_motions:].reshape(N * E, -1).abs()
            thr = torch.clamp(torch.quantile(a, q, dim=1), min=lo, max=hi)[:, None, None]
            res = torch.clamp(res, min=-thr, max=thr)
        r = [c[:, -cfg.n_motions:].clone() for c in res.chunk(E)]
        # CFG combine.  The reference accumulates in place through a VIEW of results[0]
        # (model.py:407-417), so wherever it reads results[0] it sees the running target:
        # 'independent' subtracts the already-updated entry 0; 'incremental'
        # reads results[i] for i >= 1, which are untouched.
        tgt = r[0]
        for i in range(E - 1):
            ref_i = tgt if (cfg_mode == 'independent' or i == 0) else r[i]
            if cfg_mode not in ('independent', 'incremental'):
                raise NotImplementedError(f'Unknown cfg_mode {cfg_mode}')
            tgt = tgt + cfg_scale[i] * (r[i + 1] - ref_i)
        if separate:   # the same (aliased) CFG recursion on the dynamic part, the static part and the alphas
            def combine(parts):
                parts = [c[:, -cfg.n_motions:] for c in parts]
                acc = parts[0]
                for i in range(E - 1):
                    ref_i = acc if (cfg_mode == 'independent' or i == 0) else parts[i]
                    acc = acc + cfg_scale[i] * (parts[i + 1] - ref_i)
                return acc
            tgt_dyn, tgt_stat, tgt_alpha = combine(dyn.chunk(E)), combine(stat.chunk(E)), combine(alph.chunk(E))
        if cfg.target == 'noise':
            c0 = 1 / torch.sqrt(alpha)
            c1 = (1 - alpha) / torch.sqrt(1 - ab)
            x = c0 * (x - c1 * tgt) + sigma * zt
        else:
            c0 = (1 - ab_prev) * torch.sqrt(alpha) / (1 - ab)
            c1 = (1 - alpha) * torch.sqrt(ab_prev) / (1 - ab)
            x = c0 * x + c1 * tgt + sigma * zt
        if separate:
            cum_static = cum_static + c1 * tgt_stat                     # model.py:631 / :637
            alpha_traj.append(tgt_alpha)
        traj[t - 1] = x
    if ret_traj:
        return traj, x_T, audio_feat
    if separate:
        return x, x_T, audio_feat, tgt_dyn, cum_static, torch.cat(alpha_traj, 0)
    return x, x_T, audio_feat


# ----------------------------------------------------------------------------- windowing driver
def infer_coeffs(sd, cfg, audio_feat, shape_coef, style_feat, clip_len, x_T, z_per_window, cfg_mode=None,
                 cfg_cond=None, cfg_scale=1.15, n_steps=None):
    """inference.py:34-75 from the extracted audio features on (feature extraction is oracle/audio.py).

    audio_feat [N, n_sub*L, d]; windows are sampled sequentially, each conditioned on the previous
    window's last n_prev frames of motion and of INPUT audio features; every window re-uses window 0's
    x_T (inference.py:57-69); the last window's indicator is 0 over the padded frames (:51-53) which are
    trimmed afterwards (:71-72).  z_per_window: list of [T+1, N, L, 67] noise tensors.
    """
    L, Lp = cfg.n_motions, cfg.n_prev_motions
    N, total = audio_feat.shape[:2]
    n_sub = total // L
    n_pad = total - clip_len
    prev_m = prev_a = None
    out = []
    for i in range(n_sub):
        ind = torch.ones(N, L)
        if i == n_sub - 1 and n_pad > 0:
            ind[:, -n_pad:] = 0
        a_in = audio_feat[:, i * L:(i + 1) * L]
        x0, _, used = sample(sd, cfg, a_in, shape_coef, style_feat, prev_m, prev_a, x_T, z_per_window[i], ind,
                             cfg_mode, cfg_cond, cfg_scale, n_steps=n_steps)
        prev_m = x0[:, -Lp:].clone()
        prev_a = used[:, -Lp:]
        out.append(x0[:, :-n_pad] if (i == n_sub - 1 and n_pad > 0) else x0)
    return torch.cat(out, 1)
