"""Per-kernel-class times of the audio encoder: python tools/audio_prof.py [clips] [seconds] [hubert|wav2vec2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('HF_HUB_OFFLINE', '1')
import torch, transformers
from msmd_b200 import _lib, model as M
from msmd_b200.utils import hubert, wav2vec2
from tools import synth
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
am = sys.argv[3] if len(sys.argv) > 3 else 'hubert'
args = synth.pinned_args(audio_model=am)
enc = hubert.HubertModel(transformers.HubertConfig()) if am == 'hubert' else wav2vec2.Wav2Vec2Model(transformers.Wav2Vec2Config())
m = M.MSMD(args, 'cpu', True, use_head_alpha=False, audio_encoder=enc)
m.load_state_dict(synth.fill_state_dict(synth.param_spec(m, skip=()), 1234), strict=False)
m = m.cuda().eval()
n_sub = int(secs * 25) // 100
audio = torch.stack([synth.clip_audio(i, int(secs * 16000)) for i in range(clips)]).cuda()
for _ in range(3):
    f = m.extract_audio_feature(audio, n_sub * 100)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f = m.extract_audio_feature(audio, n_sub * 100)
e1.record(); torch.cuda.synchronize()
print(f'{am} {clips} clips x {secs:g} s (encoder T = {2 * n_sub * 100}): {e0.elapsed_time(e1) / 5:.2f} ms per batch, checksum {float(f.double().abs().sum()):.4f}')
_lib.lib().msmd_profile_reset(); _lib.lib().msmd_profile_enable(1)
m.extract_audio_feature(audio, n_sub * 100)
torch.cuda.synchronize(); _lib.lib().msmd_profile_enable(0)
prof = _lib.profile_dump()
for k, (ms, n) in sorted(prof.items(), key=lambda x: -x[1][0])[:8]:
    if k and not k.startswith('gemm_') or k == 'gemm_bf16':
        print(f'   {ms:8.3f} ms  x{n:3d}  {k}')
