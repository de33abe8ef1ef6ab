"""Host-side cost of a sampling call: eager first step + graph capture/instantiate vs the replays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import SamplerWorkload
wl = SamplerWorkload(clips=64, seconds=4.0)
wl.setup(torch.device('cuda', 0), 0)
d = dict(wl.dev)
g = torch.Generator(device='cuda').manual_seed(0)
af = torch.randn(64, 100, 512, device='cuda', generator=g); st = torch.randn(64, 256, device='cuda', generator=g)
ind = torch.ones(64, 100, device='cuda')
m = wl.model
m.sample(af, d['shape'], st, motion_at_T=d['x_T'], indicator=ind, cfg_scale=1.4, noise=d['z'], n_steps=4)
eng = m._eng
def run(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.sample_window(d['x_T'], d['z'], 0, False, 1.4, 1.4, 0.0, t_start=500, n_steps=n)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for n in (3, 3, 53, 53, 500, 500):
    print(f'n_steps={n}: {run(n):.2f} ms wall')
torch.cuda.synchronize(); t0 = time.perf_counter()
x0 = m.sample(af, d['shape'], st, motion_at_T=d['x_T'], indicator=ind, cfg_scale=1.4, noise=d['z'])[0]
torch.cuda.synchronize(); print(f'MSMD.sample 500 steps (window_begin + loop): {(time.perf_counter() - t0) * 1e3:.2f} ms wall')
import cProfile, pstats
def once():
    m.sample(af, d['shape'], st, motion_at_T=d['x_T'], indicator=ind, cfg_scale=1.4, noise=d['z'], n_steps=3)
    torch.cuda.synchronize()
once()
torch.cuda.synchronize(); t0 = time.perf_counter(); once(); print(f'MSMD.sample n_steps=3: {(time.perf_counter() - t0) * 1e3:.2f} ms wall')
pr = cProfile.Profile(); pr.enable(); once(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
