"""CPU: args.json / checkpoint loading glue (inference.py:80-103, model_common.py:9-81) and the coefficient-dict
helpers (common.py:140-196) of the drop-in package."""
import argparse
import os

import torch

from helpers import make_msmd
from oracle import synth
from oracle.ref_shims import pinned_args


def test_args_round_trip(tmp_path):
    from msmd_b200.utils.model_common import load_args, load_args_with_defaults, save_args
    a = pinned_args(extra_none=None, extra_str='None')
    save_args(a, tmp_path)
    b = load_args(tmp_path)
    assert b.n_motions == 100 and b.cfg_mode == 'incremental'
    assert not hasattr(b, 'extra_none') and not hasattr(b, 'extra_str') and not hasattr(b, 'style_enc_ckpt')
    p = argparse.ArgumentParser()
    p.add_argument('--new_flag', type=int, default=7)
    p.add_argument('--n_motions', type=int, default=1)
    c = load_args_with_defaults(tmp_path, p)
    assert c.new_flag == 7 and c.n_motions == 100


def test_pretrained_audio_encoder_is_not_silently_random(monkeypatch):
    """model.py:94-101 downloads the released HuBERT / wav2vec2 weights.  Without network or cache the drop-in raises
    instead of silently returning a random encoder; the random-init fallback is an explicit opt-in and warns."""
    import pytest
    from msmd_b200.utils import hubert, wav2vec2
    monkeypatch.setenv('HF_HUB_OFFLINE', '1')
    monkeypatch.delenv('MSMD_ALLOW_RANDOM_AUDIO_ENCODER', raising=False)
    with pytest.raises(OSError):
        hubert.HubertModel.from_pretrained('facebook/hubert-base-ls960', cache_dir='/nonexistent-cache')
    with pytest.warns(RuntimeWarning, match='RANDOMLY'):
        m = wav2vec2.Wav2Vec2Model.from_pretrained('facebook/wav2vec2-base-960h', allow_random_init=True)
    assert isinstance(m, wav2vec2.Wav2Vec2Model)
    monkeypatch.setenv('MSMD_ALLOW_RANDOM_AUDIO_ENCODER', '1')
    with pytest.warns(RuntimeWarning):
        assert isinstance(hubert.HubertModel.from_pretrained('facebook/hubert-base-ls960'), hubert.HubertModel)


def test_load_model_from_reference_layout_checkpoint(tmp_path, monkeypatch):
    """A checkpoint in the reference's layout ({'model', 'style_enc', 'iter'}, training_script.py:227-233) loads
    through load_model unchanged."""
    import transformers
    monkeypatch.setenv('HF_HUB_OFFLINE', '1')
    monkeypatch.setenv('MSMD_ALLOW_RANDOM_AUDIO_ENCODER', '1')     # no HF cache here; the checkpoint overwrites every weight
    from msmd_b200.inference import load_model
    from msmd_b200.style_encoder import get_style_encoder
    from msmd_b200.utils import hubert
    from msmd_b200.utils.model_common import save_args
    from msmd_b200 import model as M
    args = pinned_args()
    exp = tmp_path / 'DPT' / 'run1'
    os.makedirs(exp / 'checkpoints')
    save_args(argparse.Namespace(**vars(args)), exp)
    src = M.MSMD(args, 'cpu', True, use_head_alpha=False, audio_encoder=hubert.HubertModel(transformers.HubertConfig()))
    src.load_state_dict(synth.fill_state_dict(synth.param_spec(src, skip=()), 5), strict=False)
    se = get_style_encoder(args, 'vae2')
    se.load_state_dict(synth.fill_state_dict(synth.param_spec(se), 6), strict=False)
    torch.save({'args': vars(args), 'model': src.state_dict(), 'style_enc': se.state_dict(), 'iter': 1000},
               exp / 'checkpoints' / 'iter_0001000.pt')
    model, style_enc, margs = load_model(str(tmp_path), 'run1', '0001000', device='cpu')
    assert margs.n_diff_steps == 500 and not model.training and not style_enc.training
    for k, v in src.state_dict().items():
        assert torch.equal(model.state_dict()[k], v), k
    for k, v in se.state_dict().items():
        assert torch.equal(style_enc.state_dict()[k], v), k


def test_get_coef_dict_layout():
    from msmd_b200.utils.common import get_coef_dict
    m = torch.randn(2, 7, 54)
    d = get_coef_dict(m, torch.randn(2, 100), with_global_pose=True)
    assert d['exp'].shape == (2, 7, 50) and d['pose'].shape == (2, 7, 6) and d['shape'].shape == (2, 7, 100)
    assert torch.equal(d['pose'][..., :4], m[..., 50:]) and (d['pose'][..., 4:] == 0).all()
    d2 = get_coef_dict(m[..., :51], with_global_pose=False)
    assert (d2['pose'][..., :3] == 0).all() and torch.equal(d2['pose'][..., 3], m[..., 50])
