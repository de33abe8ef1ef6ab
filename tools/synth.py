"""Synthetic, seeded assets / weights / inputs shared by tests, smoke and bench (input generation only: no
arithmetic of the path lives here, and nothing here reads /root/reference or imports oracle/).
Shapes and seeds follow SURVEY.md section 8(d); the pinned hyper-parameters follow SURVEY App. A.
"""
import argparse
import hashlib
import math
import numpy as np
import torch

FLAME_V = 5023
FLAME_NJ = 5
FLAME_PARENTS = [-1, 0, 1, 1, 1]


def flame_raw(seed=0, V=FLAME_V, n_total=400):
    """FLAME2020-shaped raw model (the dict the reference un-pickles, flame.py:69-90)."""
    rng = np.random.default_rng(seed)
    v_template = rng.normal(0, 0.08, (V, 3)).astype(np.float32)
    shapedirs = rng.normal(0, 1e-3, (V, 3, n_total)).astype(np.float32)
    posedirs = rng.normal(0, 1e-3, (V, 3, 36)).astype(np.float32)
    J_regressor = rng.uniform(0, 1, (FLAME_NJ, V)).astype(np.float32)
    J_regressor /= J_regressor.sum(1, keepdims=True)
    weights = rng.uniform(0, 1, (V, FLAME_NJ)).astype(np.float32)
    weights /= weights.sum(1, keepdims=True)
    faces = rng.integers(0, V, (2 * V - 70, 3)).astype(np.int64)
    kintree = np.array([[2 ** 32 - 1, 0, 1, 1, 1], [0, 1, 2, 3, 4]], dtype=np.int64)
    return dict(f=faces, v_template=v_template, shapedirs=shapedirs, posedirs=posedirs,
                J_regressor=J_regressor, kintree_table=kintree, weights=weights)


def flame_assets(seed=0, V=FLAME_V, n_shape=300, n_exp=100):
    """Buffers as FLAME.__init__ registers them (flame.py:74-90), torch fp32 on CPU."""
    raw = flame_raw(seed, V, 400)
    sd = torch.from_numpy(raw['shapedirs'])
    shapedirs = torch.cat([sd[:, :, :n_shape], sd[:, :, 300:300 + n_exp]], 2).contiguous()
    posedirs = torch.from_numpy(np.reshape(raw['posedirs'], [-1, 36]).T.copy())
    return dict(v_template=torch.from_numpy(raw['v_template']), shapedirs=shapedirs,
                posedirs=posedirs, J_regressor=torch.from_numpy(raw['J_regressor']),
                parents=list(FLAME_PARENTS), lbs_weights=torch.from_numpy(raw['weights']),
                faces=torch.from_numpy(raw['f']))


def flame_lmk_embeddings(n_faces, seed=7):
    """Landmark-embedding dict shaped like landmark_embedding.npy (flame.py:107-116):
    dynamic_* are torch tensors there (flame.py:112-113), the rest numpy."""
    rng = np.random.default_rng(seed)

    def bary(*s):
        b = rng.uniform(0.1, 1, s + (3,)).astype(np.float32)
        return b / b.sum(-1, keepdims=True)

    return dict(static_lmk_faces_idx=rng.integers(0, n_faces, 51).astype(np.int64),
                static_lmk_bary_coords=bary(51),
                dynamic_lmk_faces_idx=torch.from_numpy(rng.integers(0, n_faces, (79, 17)).astype(np.int64)),
                dynamic_lmk_bary_coords=torch.from_numpy(bary(79, 17)),
                full_lmk_faces_idx=rng.integers(0, n_faces, (1, 68)).astype(np.int64),
                full_lmk_bary_coords=bary(1, 68))


def flame_inputs(B, n_shape=300, n_exp=100, seed=0):
    """Config-2 inputs: shape, exp ~N(0,1); pose 0.2 N; eye 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    shape = torch.randn(B, n_shape, generator=g)
    exp = torch.randn(B, n_exp, generator=g)
    pose = 0.2 * torch.randn(B, 6, generator=g)
    eye = 0.1 * torch.randn(B, 6, generator=g)
    return shape, exp, pose, eye


def _key_seed(name, seed):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return int.from_bytes(h[:8], 'little') % (2 ** 63 - 1)


def fill_state_dict(spec, seed=1234):
    """Deterministic weights for a {name: shape} spec, independent of module init order.

    Linear/conv weights ~ U(+-1/sqrt(fan_in)); biases small; norm gains 1+0.1 N, norm
    biases 0.1 N; free parameters (PE, start/null feats) ~ N(0,1) like the reference
    constructors (model.py:117-137, :864).  Buffers (TE.pe, schedule, masks) are NOT
    produced here - modules build them themselves.
    """
    out = {}
    for name, shape in spec.items():
        g = torch.Generator().manual_seed(_key_seed(name, seed))
        shape = tuple(shape)
        leaf = name.split('.')[-1]
        is_norm = ('norm' in name.lower()) or name.endswith(('input_layers.5.weight', 'input_layers.5.bias',
                                                              'input_layers.11.weight', 'input_layers.11.bias',
                                                              'output_layers.5.weight', 'output_layers.5.bias'))
        if is_norm and leaf == 'weight' and len(shape) == 1:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_norm and leaf == 'bias':
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == 'masked_spec_embed':
            t = torch.rand(shape, generator=g)
        elif leaf in ('weight', 'in_proj_weight', 'original1', 'weight_v') and len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif leaf in ('original0', 'weight_g'):
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf in ('bias', 'in_proj_bias'):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g)
        out[name] = t.float()
    return out


def clip_audio(clip_id, n_samples=160000):
    """4 sines + noise, then (x-mean)/(std+1e-5) (inference.py:234)."""
    g = torch.Generator().manual_seed(1000 + clip_id)
    t = torch.arange(n_samples, dtype=torch.float64) / 16000.0
    x = sum(torch.sin(2 * math.pi * f * t) for f in (110.0, 220.0, 440.0, 880.0)) * 0.5
    x = x.float() + 0.1 * torch.randn(n_samples, generator=g)
    return (x - x.mean()) / (x.std() + 1e-5)


def clip_style_motion(clip_id, L=100, d=67):
    return torch.randn(1, L, d, generator=torch.Generator().manual_seed(2000 + clip_id))


def clip_style_eps(clip_id, d_style=256):
    return torch.randn(1, d_style, generator=torch.Generator().manual_seed(3000 + clip_id))


def clip_xT(clip_id, L=100, d=67):
    return torch.randn(1, L, d, generator=torch.Generator().manual_seed(4000 + clip_id))


def clip_step_noise(clip_id, window, n_steps, L=100, d=67):
    """z[t] for t = n_steps..1 stacked as [n_steps+1, L, d]; index t; z[1] = z[0] = 0 (model.py:378-381)."""
    g = torch.Generator().manual_seed(_key_seed(f"z:{clip_id}:{window}", 5000))
    z = torch.randn(n_steps + 1, L, d, generator=g)
    z[0] = 0
    z[1] = 0
    return z


def param_spec(module, skip=('audio_encoder.',)):
    """{name: shape} of a module's trainable parameters (buffers are rebuilt by the modules)."""
    return {k: tuple(v.shape) for k, v in module.named_parameters() if not k.startswith(skip)}


def denoiser_inputs(N, seed=0, L=100, Lp=10, d=512, dm=67, d_style=256, T=500):
    """Seeded inputs of one DenoisingNetwork_MSMD.forward call (model.py:914)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    ind = torch.ones(N, L)
    if N > 1:
        ind[1, -30:] = 0
    step = torch.randint(1, T + 1, (N,), generator=g)
    return dict(motion=r(N, L, dm), audio=r(N, L, d), person=r(N, 1, 100 + d_style), style=r(N, 1, d_style),
                prev_motion=r(N, Lp, dm), prev_audio=r(N, Lp, d), step=step, indicator=ind)


def sampler_inputs(N, T, seed=0, L=100, d=512, dm=67, d_style=256):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    z = r(T + 1, N, L, dm)
    z[0] = 0
    z[1] = 0
    ind = torch.ones(N, L)
    ind[-1, -17:] = 0
    return dict(audio_feat=r(N, L, d), shape=torch.zeros(N, 1, 100), style=r(N, d_style), x_T=r(N, L, dm), z=z,
                indicator=ind)


def pinned_args(**over):
    """SURVEY App. A pinned configuration (the hyper-parameters model.py reads)."""
    a = dict(audio_model='hubert', style_enc_model_style='vae2', d_style=256, num_of_basis=4,
             use_indicator=True, n_motions=100, n_prev_motions=10, fps=25,
             architecture='decoder', feature_dim=512, n_heads=8, n_layers=8, mlp_ratio=4,
             align_mask_width=1, no_use_learnable_pe=False, n_diff_steps=500,
             diff_schedule='cosine', target='sample', cfg_mode='incremental',
             guiding_conditions='audio,style', style_enc_ckpt=None, regularize_alpha='None',
             dataset_type='ravdess+celebv-text-medium', rot_repr='euler', no_head_pose=False)
    a.update(over)
    return argparse.Namespace(**a)
