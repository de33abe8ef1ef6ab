"""Warm-cache per-kernel-class breakdown of sampling steps (CUDA events around every launch, eager mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import SamplerWorkload
from msmd_b200 import _lib
clips = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wl = SamplerWorkload(clips=clips, seconds=4.0)
wl.precision = os.environ.get("MSMD_PRECISION", "bf16")
wl.setup(torch.device('cuda', 0), 0)
d = dict(wl.dev)
g = torch.Generator(device='cuda').manual_seed(0)
af = torch.randn(clips, 100, 512, device='cuda', generator=g)
st = torch.randn(clips, 256, device='cuda', generator=g)
ind = torch.ones(clips, 100, device='cuda')
m = wl.model
if os.environ.get('MSMD_PRECISION'):
    m.precision = m.denoising_net.precision = os.environ['MSMD_PRECISION']
m.sample(af, d['shape'], st, motion_at_T=d['x_T'], indicator=ind, cfg_scale=1.4, noise=d['z'], n_steps=4)
eng = m._eng
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.sample_window(d['x_T'], d['z'], 0, False, 1.4, 1.4, 0.0, t_start=500, n_steps=200); e1.record()
torch.cuda.synchronize()
print(f'graph replay: {e0.elapsed_time(e1) / 200 * 1000:.1f} us per sampling step (200 steps incl. 1 eager + capture)')
_lib.lib().msmd_profile_reset(); _lib.lib().msmd_profile_enable(1)
eng.sample_window(d['x_T'], d['z'], 0, False, 1.4, 1.4, 0.0, t_start=500, n_steps=steps)
torch.cuda.synchronize()
_lib.lib().msmd_profile_enable(0)
prof = _lib.profile_dump()
tot = sum(v[0] for k, v in prof.items() if k and (not k.startswith('gemm_') or k == 'gemm_bf16'))
for k, (ms, n) in sorted(prof.items(), key=lambda x: -x[1][0]):
    if not k: continue
    print(f'{ms / steps * 1000:9.1f} us/step  x{n // steps:3d}  {ms / n * 1000:7.1f} us each  {k}')
print(f'sum of kernel classes: {tot / steps * 1000:.1f} us/step')
