"""Experiment: the 64-clip sampling step as ONE engine / one stream vs. K engines of 64/K clips on K streams (clips are
independent, so the HBM-bound kernels and the launch ramp / tail of one sub-batch can run under the other's GEMMs).
python tools/two_stream.py [clips] [K ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import SamplerWorkload
from msmd_b200._engine import DenoiserEngine
CL = int(sys.argv[1]) if len(sys.argv) > 1 else 64
KS = [int(a) for a in sys.argv[2:]] or [1, 2]
wl = SamplerWorkload(clips=CL, seconds=4.0)
wl.precision = os.environ.get("MSMD_PRECISION", "bf16")
wl.setup(torch.device('cuda', 0), 0)
d = dict(wl.dev)
g = torch.Generator(device='cuda').manual_seed(0)
af = torch.randn(CL, 100, 512, device='cuda', generator=g)
st = torch.randn(CL, 256, device='cuda', generator=g)
ind = torch.ones(CL, 100, device='cuda')
m = wl.model
m.sample(af, d['shape'], st, motion_at_T=d['x_T'], indicator=ind, cfg_scale=1.4, noise=d['z'], n_steps=4)
cfg_fn, sd_fn = m._engine_state()
sd = sd_fn()
N_STEPS = 150
ref = None
for K in KS:
    n = CL // K
    engs, streams = [], [torch.cuda.Stream() for _ in range(K)]
    for k in range(K):
        e = DenoiserEngine(cfg_fn(3 * n), torch.device('cuda', 0))
        e.load_state_dict(sd)
        engs.append(e)
    torch.cuda.synchronize()
    # open the windows through the model (conditioning of the 3 CFG entries), one engine at a time
    outs = [None] * K
    zs = [d['z'][:, k * n:(k + 1) * n].contiguous() for k in range(K)]
    def run(steps):
        for k in range(K):
            with torch.cuda.stream(streams[k]):
                sl = slice(k * n, (k + 1) * n)
                outs[k] = engs[k].sample_window(d['x_T'][sl], zs[k], 0, False, 1.4, 1.4, 0.0,
                                                t_start=500, n_steps=steps)[0]
    for k in range(K):
        sl = slice(k * n, (k + 1) * n)
        object.__setattr__(m, '_eng', engs[k]); engs[k].weights_key = tuple((kk, v.data_ptr(), v._version) for kk, v in sd.items())
        with torch.cuda.stream(streams[k]):
            m.sample(af[sl], d['shape'][sl], st[sl], motion_at_T=d['x_T'][sl], indicator=ind[sl], cfg_scale=1.4,
                     noise=zs[k], n_steps=4)
    torch.cuda.synchronize()
    run(8); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s_ in streams: s_.wait_stream(torch.cuda.current_stream())
        run(N_STEPS)
        for s_ in streams: torch.cuda.current_stream().wait_stream(s_)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / N_STEPS * 1000)
    x = torch.cat(outs, 0)
    if ref is None: ref = x
    print(f'K={K}: {best:.1f} us per sampling step of {CL} clips (best of 3 x {N_STEPS} replays); max |x - x(K={KS[0]})| = {float((x - ref).abs().max()):.3e}')
    del engs
