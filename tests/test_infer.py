"""inference.infer_coeffs (config 1: one 10 s clip -> HuBERT -> 3 windows, last one padded):
oracle vs the golden vector produced by the reference's own driver (CPU); CUDA drop-in vs golden (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import audio as A, denoiser as D, synth
from oracle.make_golden import INFER_GOLD, infer_inputs
from oracle.ref_shims import pinned_args


def build(device='cpu'):
    import transformers
    from msmd_b200 import model as M
    from msmd_b200.utils import hubert
    args = pinned_args(n_diff_steps=INFER_GOLD['T'])
    m = M.MSMD(args, 'cpu', True, use_head_alpha=False, audio_encoder=hubert.HubertModel(transformers.HubertConfig()))
    m.load_state_dict(synth.fill_state_dict(synth.param_spec(m, skip=()), INFER_GOLD['weight_seed']), strict=False)
    return m.to(device).eval(), args


def test_oracle_infer_coeffs_matches_reference_driver():
    m, args = build()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = np.load(os.path.join(GOLDEN, 'infer.npz'))
    audio, style, x_T, z = infer_inputs()
    padded = torch.nn.functional.pad(audio, (0, 3 * 64000 - audio.numel())).unsqueeze(0)
    feat = A.extract_audio_feature(sd, padded, 25, 300)
    assert rel_l2(feat, g['audio_feat']) < 5e-6
    out = D.infer_coeffs(sd, args, feat, torch.zeros(1, 1, 100), style, 250, x_T, z, cfg_scale=1.4)
    assert out.shape == (1, 250, 67) and rel_l2(out, g['coeffs']) < 2e-5


@pytest.mark.gpu
def test_infer_coeffs_cuda_matches_golden(built_lib):
    from msmd_b200.inference import infer_coeffs, infer_coeffs_batched
    m, args = build('cuda')
    g = np.load(os.path.join(GOLDEN, 'infer.npz'))
    audio, style, x_T, z = infer_inputs()
    padded = torch.nn.functional.pad(audio, (0, 3 * 64000 - audio.numel())).unsqueeze(0).cuda()
    feat = m.extract_audio_feature(padded, 300)
    e_feat = rel_l2(feat, g['audio_feat'])
    out = infer_coeffs_batched(m, args, feat, torch.zeros(1, 1, 100).cuda(), style.cuda(), clip_len=250, cfg_scale=1.4,
                               x_T=x_T.cuda(), noise=[zz.cuda() for zz in z])
    e_out = rel_l2(out, g['coeffs'])
    print('infer_coeffs: audio feature rel-L2', e_feat, ' coefficients rel-L2 (6 free-running bf16 steps x 3 windows)', e_out)
    assert out.shape == (1, 250, 67) and e_feat < 2e-2 and e_out < 5e-2
    # the reference-signature entry point (draws its own x_T / noise): shape, trimming, determinism under a seed
    torch.manual_seed(11)
    a = infer_coeffs(m, args, audio.cuda(), torch.zeros(1, 1, 100).cuda(), 640.0, style.cuda(), cfg_scale=1.4,
                     dynamic_threshold=None)
    torch.manual_seed(11)
    b = infer_coeffs(m, args, audio.cuda(), torch.zeros(1, 1, 100).cuda(), 640.0, style.cuda(), cfg_scale=1.4,
                     dynamic_threshold=None)
    assert a.shape == (1, 250, 67) and torch.equal(a, b) and torch.isfinite(a).all()
    short = infer_coeffs(m, args, audio[:48000].cuda(), torch.zeros(1, 1, 100).cuda(), 640.0, style.cuda(),
                         n_repetitions=2, cfg_scale=1.4, dynamic_threshold=None)
    assert short.shape == (2, 75, 67)
