"""Engine-level behaviour behind the drop-in modules: keep_separate outputs, step-graph reuse without hidden
synchronisation (include/msmd_b200.h conventions), weight re-packing without leaks, argument validation, and parity at
the BENCHMARK size (64 clips x 3 CFG entries = 192 sequences: CTA-pair GEMMs, the 148-CTA attention grid, 0.5 GB caches)."""
import os
import time

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from helpers import cpu_state_dict, make_msmd
from oracle import denoiser as D, synth
from oracle.make_golden import DEN_GOLD

pytestmark = pytest.mark.gpu
F32_TOL = 1e-5


def test_keep_separate_matches_reference_golden(built_lib):
    """DenoisingNetwork_MSMD.forward(keep_separate=True) -> (dynamic, static tiled over rows, alphas), model.py:972-973."""
    g = np.load(os.path.join(GOLDEN, 'denoiser_parts.npz'))
    m, args = make_msmd('cuda', precision=None)
    i = {k: v.cuda() for k, v in synth.denoiser_inputs(DEN_GOLD['N'], DEN_GOLD['seed']).items()}
    call = lambda **kw: m.denoising_net(i['motion'], i['audio'], i['person'], i['style'], i['prev_motion'],
                                        i['prev_audio'], i['step'], i['indicator'], keep_separate=True, **kw)
    dyn, sta, alp = call()
    assert dyn.shape == g['dyn'].shape and sta.shape == g['static'].shape and alp.shape == g['alphas'].shape
    e = [rel_l2(dyn, g['dyn']), rel_l2(sta, g['static']), rel_l2(alp, g['alphas'])]
    print('keep_separate rel-L2 (dyn, static, alphas) vs reference golden:', e)
    assert max(e) < F32_TOL
    d16, s16, a16 = call(precise='bf16')
    assert rel_l2(d16, g['dyn']) < 1.5e-2 and rel_l2(a16, g['alphas']) < 3e-2 and torch.equal(s16, sta)
    # the parts recombine to the mixed output (model.py:985-995)
    mixed = m.denoising_net(i['motion'], i['audio'], i['person'], i['style'], i['prev_motion'], i['prev_audio'], i['step'],
                            i['indicator'])
    face = (sta[..., :-3] * alp.unsqueeze(-1)).sum(2)
    pose = sta[..., -3:].sum(2)
    assert rel_l2(dyn + torch.cat([face, pose], -1), mixed) < 1e-6


def test_step_graph_is_cached_and_the_call_does_not_synchronise(built_lib):
    """The second window of a shape launches the graph instantiated by the first (no capture / instantiate / sync):
    the call returns long before the stream drains, results are identical, and later windows with OTHER caller buffers
    (noise, x_T, scales) reuse the same graph correctly."""
    T, N = 500, 8
    m, args = make_msmd('cuda', n_diff_steps=T)
    i = {k: v.cuda() for k, v in synth.sampler_inputs(N, T, 3).items()}
    run = lambda x_T, z, scale: m.sample(i['audio_feat'], i['shape'], i['style'], motion_at_T=x_T, indicator=i['indicator'],
                                         cfg_scale=scale, noise=z)[0]
    a = run(i['x_T'], i['z'], 1.4)                      # first call: one eager step + capture + replays
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    b = run(i['x_T'], i['z'], 1.4)
    host_ms = (time.perf_counter() - t0) * 1e3
    ev1.record()
    drained_at_return = ev1.query()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    print(f'500-step window, {3 * N} sequences: host returned after {host_ms:.1f} ms, device busy {dev_ms:.1f} ms')
    assert not drained_at_return and host_ms < 0.6 * dev_ms
    assert torch.equal(a, b)
    # fresh caller tensors (different addresses and values) through the cached graph == a fresh engine
    z2, x2 = (i['z'] * 0.5).clone(), (i['x_T'] + 0.1).clone()
    c = run(x2, z2, 1.7)
    m2, _ = make_msmd('cuda', n_diff_steps=T)
    c_ref = m2.sample(i['audio_feat'], i['shape'], i['style'], motion_at_T=x2, indicator=i['indicator'], cfg_scale=1.7,
                      noise=z2)[0]
    assert torch.equal(c, c_ref) and not torch.equal(c, a)
    # conditioning tensors may be freed right after window_begin: the engine keeps its own copies
    eng = m._eng
    assert len(getattr(eng, '_keep', None) or ()) == 0


def test_weight_reload_replaces_packed_copies(built_lib):
    """A parameter update re-packs the weights in place of the old copies (they used to accumulate until destroy), and a
    standalone module forward does not re-pack at all when nothing changed."""
    m, args = make_msmd('cuda', precision=None)
    net = m.denoising_net
    i = {k: v.cuda() for k, v in synth.denoiser_inputs(2, 7).items()}
    call = lambda: net(i['motion'], i['audio'], i['person'], i['style'], i['prev_motion'], i['prev_audio'], i['step'],
                       i['indicator'])
    a = call()
    key = net._eng.weights_key
    assert torch.equal(call(), a) and net._eng.weights_key == key        # stable key: no reload between identical calls
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for r in range(4):
        with torch.no_grad():
            net.motion_dec[2].bias[:67].add_(0.01)                         # dynamic-part biases; bumps the parameter version -> reload
        b = call()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert net._eng.weights_key != key and not torch.equal(a, b)
    assert abs((b - a).mean().item() - 0.04) < 1e-4                      # the dynamic part moved by the bias delta
    assert free0 - free1 < 32 * 2 ** 20, f'weight reload leaked {(free0 - free1) / 2 ** 20:.0f} MiB'


def test_step_indices_and_shapes_are_validated(built_lib):
    m, args = make_msmd('cuda', precision=None)
    i = {k: v.cuda() for k, v in synth.denoiser_inputs(1, 3).items()}
    f = lambda step: m.denoising_net(i['motion'], i['audio'], i['person'], i['style'], i['prev_motion'], i['prev_audio'],
                                     step, i['indicator'])
    for bad in (torch.tensor([501]), torch.tensor([-1]), 777):
        with pytest.raises(IndexError, match='diffusion step'):
            f(bad)
    assert torch.isfinite(f(torch.tensor([500]))).all() and torch.isfinite(f(0)).all()
    assert torch.isfinite(f(torch.tensor([9999], device='cuda'))).all()     # device-side indices are clamped, never OOB
    from msmd_b200._lib import MsmdError
    with pytest.raises((MsmdError, ValueError), match='112-token'):
        make_msmd('cuda', n_motions=750, n_prev_motions=100)[0].denoising_net(
            torch.zeros(1, 750, 67).cuda(), torch.zeros(1, 750, 512).cuda(), i['person'], i['style'],
            torch.zeros(1, 100, 67).cuda(), torch.zeros(1, 100, 512).cuda(), 3, torch.ones(1, 750).cuda())


def test_fp32_grade_path_with_narrow_ffn(built_lib):
    """mlp_ratio = 1 (d_ff < 2 d_model): the fp32-grade window pass splits the kv cache [S*Tk, 2d] into the operand
    scratch, which used to be sized for [M, d_ff] only."""
    m, args = make_msmd('cuda', precision='fp32', mlp_ratio=1)
    i = synth.denoiser_inputs(3, 9)
    want = D.denoiser_forward(cpu_state_dict(m), args, i['motion'], i['audio'], i['person'], i['style'],
                              i['prev_motion'], i['prev_audio'], i['step'], i['indicator'])
    c = {k: v.cuda() for k, v in i.items()}
    got = m.denoising_net(c['motion'], c['audio'], c['person'], c['style'], c['prev_motion'], c['prev_audio'], c['step'],
                          c['indicator'])
    assert rel_l2(got, want) < F32_TOL


# ------------------------------------------------------------------------------------------ benchmark-size parity
BENCH_CLIPS = 64
CHECK = (0, 31, 63)


def _bench_inputs(T):
    i = synth.sampler_inputs(BENCH_CLIPS, T, 17)
    i['indicator'][5, -40:] = 0
    return i


def _sub(i, idx):
    ix = torch.tensor(idx)
    return dict(audio_feat=i['audio_feat'][ix], shape=i['shape'][ix], style=i['style'][ix], x_T=i['x_T'][ix],
                z=i['z'][:, ix], indicator=i['indicator'][ix])


def test_benchmark_size_teacher_forced_steps_vs_oracle(built_lib):
    """configs[2] shape: 64 clips -> 192 sequences, M = 21312 rows.  Teacher-forced single steps at both ends of every
    segment of the default schedule; clips 0 / 31 / 63 against the CPU oracle run on those clips ALONE (clips are
    independent, so this is exact).  Flat 1e-3 for the default (hybrid) engine, 1e-5 for the fp32-grade one."""
    T = 500
    i = _bench_inputs(T)
    sub = _sub(i, CHECK)
    dev = {k: v.cuda() for k, v in i.items()}
    mh, args = make_msmd('cuda', precision=None, n_diff_steps=T)
    m32, _ = make_msmd('cuda', precision='fp32', n_diff_steps=T)
    sd = cpu_state_dict(mh)
    k32, k16 = mh._precise_steps(), mh._fp16_steps()
    kw = dict(cfg_mode='incremental', cfg_scale=[1.4, 1.4])
    for t in (500, 250, k16 + 1, k16, k32 + 1, k32, 1):
        scale = 0.3 + 0.7 * t / T
        want, _, _ = D.sample(sd, args, sub['audio_feat'], sub['shape'], sub['style'], x_T=sub['x_T'] * scale, z=sub['z'],
                              indicator=sub['indicator'], t_start=t, n_steps=1, **kw)
        for name, m, tol in (('hybrid', mh, 1e-3), ('fp32', m32, F32_TOL)):
            got, _, _ = m.sample(dev['audio_feat'], dev['shape'], dev['style'], motion_at_T=dev['x_T'] * scale,
                                 indicator=dev['indicator'], noise=dev['z'], t_start=t, n_steps=1, **kw)
            errs = [rel_l2(got[c], want[j]) for j, c in enumerate(CHECK)]
            print(f't={t:3d} {name:6s} rel-L2 of clips {CHECK}: ' + ' '.join(f'{e:.2e}' for e in errs))
            assert max(errs) < tol, (t, name, errs)
    mh.check()


def test_benchmark_size_full_window_vs_oracle(built_lib):
    """One full free-running window (all T steps, CUDA-graph replays + the schedule's fp16 / fp32-grade tail) at 192
    sequences: clips 0 / 31 / 63 against the oracle sampling those clips alone.  Free-running trajectories diverge from
    the fp32 one at the rate the 16-bit steps inject error, so the hybrid bound is loose (the per-step contract is the
    teacher-forced test) - but the last, fp32-grade steps pull the state back: x_0 = x0_hat(x_1) exactly, and the
    network is insensitive to the small error x_1 carries, so the FINAL codes agree to 1e-3 (measured 1e-5).  The
    fp32-grade engine tracks the oracle over the whole window.  The CPU oracle needs ~2 minutes for the 500 steps."""
    T = 500
    i = _bench_inputs(T)
    sub = _sub(i, CHECK)
    dev = {k: v.cuda() for k, v in i.items()}
    mh, args = make_msmd('cuda', precision=None, n_diff_steps=T)
    m32, _ = make_msmd('cuda', precision='fp32', n_diff_steps=T)
    kw = dict(cfg_mode='incremental', cfg_scale=[1.4, 1.4])
    want, _, _ = D.sample(cpu_state_dict(mh), args, sub['audio_feat'], sub['shape'], sub['style'], x_T=sub['x_T'],
                          z=sub['z'], indicator=sub['indicator'], **kw)
    for name, m, tol in (('hybrid', mh, 1e-3), ('fp32', m32, 1e-5)):
        got, _, _ = m.sample(dev['audio_feat'], dev['shape'], dev['style'], motion_at_T=dev['x_T'],
                             indicator=dev['indicator'], noise=dev['z'], **kw)
        errs = [rel_l2(got[c], want[j]) for j, c in enumerate(CHECK)]
        print(f'full {T}-step window, {name}: rel-L2 of clips {CHECK}: ' + ' '.join(f'{e:.2e}' for e in errs))
        assert max(errs) < tol, (name, errs)
        # the same clips sampled alone on the same engine give the same codes (batch-size independence of every kernel)
        alone, _, _ = m.sample(dev['audio_feat'][[31]], dev['shape'][[31]], dev['style'][[31]], motion_at_T=dev['x_T'][[31]],
                               indicator=dev['indicator'][[31]], noise=dev['z'][:, [31]].contiguous(), **kw)
        assert rel_l2(alone[0], got[31]) < (1e-6 if name == 'fp32' else 1e-3)
