"""Style encoder: oracle vs golden (CPU), CUDA drop-in vs golden / oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import style as S, synth
from oracle.make_golden import STYLE_GOLD, style_inputs
from oracle.ref_shims import pinned_args


def make_style(device='cpu'):
    from msmd_b200.style_encoder import get_style_encoder
    enc = get_style_encoder(pinned_args(), 'vae2').eval()
    enc.load_state_dict(synth.fill_state_dict(synth.param_spec(enc), STYLE_GOLD['weight_seed']), strict=False)
    return enc.to(device)


def test_oracle_style_matches_golden():
    enc = make_style()
    sd = {k: v.detach() for k, v in enc.state_dict().items()}
    g = np.load(os.path.join(GOLDEN, 'style.npz'))
    x, eps, x2 = style_inputs()
    out, mu, logvar = S.style_forward(sd, x, eps)
    assert rel_l2(out, g['out']) < 2e-6 and rel_l2(mu, g['mu']) < 2e-6 and rel_l2(logvar, g['logvar']) < 2e-6
    mu2, lv2 = S.style_stats(sd, x2)
    assert rel_l2(mu2, g['mu_short']) < 2e-6 and rel_l2(lv2, g['logvar_short']) < 2e-6


def test_style_state_dict_layout():
    enc = make_style()
    want = {'input_layers.1.weight': (512, 67, 3), 'input_layers.5.weight': (512,), 'input_layers.7.weight': (512, 512, 3),
            'input_layers.11.bias': (512,), 'PE.pe': (1, 600, 512), 'encoder.self_attn.in_proj_weight': (1536, 512),
            'encoder.linear1.weight': (512, 512), 'encoder.norm2.weight': (512,), 'output_layers.1.weight': (512, 512, 3),
            'output_layers.5.bias': (512,), 'output_layers.7.weight': (512, 512, 3)}
    sd = enc.state_dict()
    assert len(sd) == 27
    for k, shp in want.items():
        assert tuple(sd[k].shape) == shp, k


@pytest.mark.gpu
def test_style_cuda_matches_golden(built_lib):
    """bf16 tensor-core GEMMs: mu / logvar within 1e-2 relative L2 of the fp32 reference."""
    enc = make_style('cuda')
    g = np.load(os.path.join(GOLDEN, 'style.npz'))
    x, eps, x2 = style_inputs()
    mu, logvar = enc._stats(x.cuda())
    e_mu, e_lv = rel_l2(mu, g['mu']), rel_l2(logvar, g['logvar'])
    print('style mu/logvar rel-L2 (bf16 vs fp32 reference):', e_mu, e_lv)
    assert e_mu < 1e-2 and e_lv < 1e-2
    mu2, lv2 = enc._stats(x2.cuda())                       # shorter clip, fewer rows than a tile
    assert rel_l2(mu2, g['mu_short']) < 1e-2 and rel_l2(lv2, g['logvar_short']) < 1e-2
    # forward()/sample() noise semantics: torch's generator, eps drawn once / twice
    torch.manual_seed(3)
    out, mu_f, lv_f = enc(x.cuda())
    torch.manual_seed(3)
    e1 = torch.randn_like(mu_f)
    assert torch.allclose(out, mu_f + e1 * torch.exp(0.5 * lv_f))
    torch.manual_seed(3)
    smp = enc.sample(x.cuda())
    torch.manual_seed(3)
    torch.randn_like(mu_f)
    e2 = torch.randn_like(mu_f)
    assert torch.allclose(smp, mu_f + e2 * torch.exp(0.5 * lv_f))
    # clips are independent
    mu1, _ = enc._stats(x[:1].cuda())
    assert rel_l2(mu1, mu[:1]) < 1e-6
