"""Short sampling run (config-3 shapes) for ncu launch lists: python tools/sampler_short.py [clips] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import SamplerWorkload
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = SamplerWorkload(clips=clips, seconds=4.0)
wl.setup(torch.device('cuda', 0), 0)
d = dict(wl.dev)
d['audio_feat'] = torch.randn(clips, 100, 512, device='cuda', generator=torch.Generator(device='cuda').manual_seed(0))
d['style'] = torch.randn(clips, 256, device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))
m = wl.model
ind = torch.ones(clips, 100, device='cuda')
for _ in range(2):
    x, _, _ = m.sample(d['audio_feat'][:, :100], d['shape'], d['style'], motion_at_T=d['x_T'], indicator=ind,
                       cfg_scale=1.4, noise=d['z'], n_steps=steps)
torch.cuda.synchronize()
print('ok', float(x.abs().mean()))
