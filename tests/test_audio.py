"""Audio encoder (HuBERT / wav2vec2 base + extract_audio_feature): oracle vs golden (CPU),
CUDA drop-in vs golden (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import audio as A, synth
from oracle.make_golden import AUDIO_GOLD, audio_inputs
from oracle.ref_shims import pinned_args


def make_msmd_with_audio(audio_model, device='cpu'):
    import transformers
    from msmd_b200 import model as M
    from msmd_b200.utils import hubert, wav2vec2
    if audio_model == 'hubert':
        enc = hubert.HubertModel(transformers.HubertConfig())
    else:
        enc = wav2vec2.Wav2Vec2Model(transformers.Wav2Vec2Config())
    m = M.MSMD(pinned_args(audio_model=audio_model), 'cpu', True, use_head_alpha=False, audio_encoder=enc)
    fill = synth.fill_state_dict(synth.param_spec(m, skip=('denoising_net.',)), AUDIO_GOLD['weight_seed'])
    missing, unexpected = m.load_state_dict(fill, strict=False)
    assert not unexpected
    return m.to(device).eval()


@pytest.mark.parametrize('am', ['hubert', 'wav2vec2'])
def test_oracle_audio_matches_golden(am):
    m = make_msmd_with_audio(am)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = np.load(os.path.join(GOLDEN, 'audio.npz'))
    x, xs = audio_inputs()
    c = AUDIO_GOLD
    assert rel_l2(A.extract_audio_feature(sd, x, 25, c['frames']), g[am]) < 5e-6
    assert rel_l2(A.extract_audio_feature(sd, xs, 25, c['short_frames']), g[am + '_short']) < 5e-6


def test_pad_audio_matches_host_helper():
    from msmd_b200.utils.model_common import pad_audio
    for n in (64000, 48000, 160000, 16001, 12345, 192000):
        x = torch.randn(2, n)
        assert torch.equal(pad_audio(x), A.pad_audio(x))
    assert A.pad_audio(torch.randn(1, 192000)).shape[1] == 192080


@pytest.mark.gpu
@pytest.mark.parametrize('am', ['hubert', 'wav2vec2'])
def test_audio_cuda_matches_golden(built_lib, am):
    """bf16 tensor-core GEMMs through 7 convs + 12 encoder layers: features within 2e-2 relative L2."""
    m = make_msmd_with_audio(am, 'cuda')
    g = np.load(os.path.join(GOLDEN, 'audio.npz'))
    x, xs = audio_inputs()
    c = AUDIO_GOLD
    got = m.extract_audio_feature(x.cuda(), c['frames'])
    err = rel_l2(got, g[am])
    print(am, 'audio feature rel-L2 (bf16 vs fp32 reference):', err)
    assert got.shape == g[am].shape and err < 2e-2
    got_s = m.extract_audio_feature(xs.cuda(), c['short_frames'])
    assert rel_l2(got_s, g[am + '_short']) < 2e-2
    one = m.extract_audio_feature(x[:1].cuda(), c['frames'])      # clips are independent (GroupNorm is per clip)
    assert rel_l2(one, got[:1]) < 1e-5


@pytest.mark.gpu
def test_audio_cuda_long_clip_wav2vec2(built_lib):
    """BASELINE configs[4]: wav2vec2 on a 60 s clip -> 3000 encoder frames (the flash-attention path of the audio
    encoder, 12 heads x 3000 x 3000 scores never materialised) vs the CPU oracle on the same clip."""
    m = make_msmd_with_audio('wav2vec2', 'cuda')
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    n = 60 * 16000
    x = torch.stack([synth.clip_audio(7, n)])
    frames = 60 * 25
    want = A.extract_audio_feature(sd, x, 25, frames)
    got = m.extract_audio_feature(x.cuda(), frames)
    err = rel_l2(got, want)
    print('wav2vec2 60 s clip: audio feature rel-L2 (bf16 vs fp32 oracle):', err)
    assert got.shape == (1, frames, 512) and err < 2e-2
