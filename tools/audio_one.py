import sys, os, torch, transformers
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msmd_b200 import model as M
from msmd_b200.utils import hubert
from tools import synth
from tools.synth import pinned_args
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = M.MSMD(pinned_args(), 'cpu', True, use_head_alpha=False, audio_encoder=hubert.HubertModel(transformers.HubertConfig()))
m.load_state_dict(synth.fill_state_dict(synth.param_spec(m, skip=('denoising_net.',)), 1), strict=False)
m = m.cuda().eval()
x = torch.stack([synth.clip_audio(i, 192000) for i in range(N)]).cuda()
for _ in range(2): f = m.extract_audio_feature(x, 300)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); f = m.extract_audio_feature(x, 300); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'audio encoder: {N} clips x 12 s in {ms:.2f} ms = {N * 12 / ms * 1e3:.0f} audio-s/s; ~{N * 12 * 15.0e9 / ms / 1e9:.0f} TFLOP/s (15 GFLOP per audio-second)')
