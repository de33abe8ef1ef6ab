"""GPU parity of the denoiser / sampler (bf16 tensor-core mode) against the fp32 reference golden
vectors and the CPU oracle.

Tolerances (BASELINE.json north_star): in bf16 the network output x0_hat carries bf16 operand rounding
through 8 layers; SURVEY App. D-5 measures 6.4e-3 relative L2 for the reference's own bf16 autocast.
We bound x0_hat at 1.5e-2 and the teacher-forced sampling step x_{t-1} = c0 x_t + c1 x0_hat + sigma z at
1e-3 + 1.5e-2 * c1(t) (the c1-damped propagation of that same error)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from helpers import cpu_state_dict, make_msmd
from oracle import denoiser as D, synth
from oracle.make_golden import DEN_GOLD, SAMP_GOLD

pytestmark = pytest.mark.gpu
X0_TOL = 1.5e-2


def test_denoiser_forward_matches_golden(built_lib):
    m, args = make_msmd('cuda')
    i = {k: v.cuda() for k, v in synth.denoiser_inputs(DEN_GOLD['N'], DEN_GOLD['seed']).items()}
    got = m.denoising_net(i['motion'], i['audio'], i['person'], i['style'], i['prev_motion'], i['prev_audio'],
                          i['step'], i['indicator'])
    want = np.load(os.path.join(GOLDEN, 'denoiser.npz'))['out']
    err = rel_l2(got, want)
    print('denoiser x0_hat rel-L2 (bf16 vs fp32 reference):', err)
    assert got.shape == want.shape and err < X0_TOL


@pytest.mark.parametrize('N', [1, 5])
def test_denoiser_forward_vs_oracle_other_batches(built_lib, N):
    m, args = make_msmd('cuda')
    i = synth.denoiser_inputs(N, seed=100 + N)
    want = D.denoiser_forward(cpu_state_dict(m), args, i['motion'], i['audio'], i['person'], i['style'],
                              i['prev_motion'], i['prev_audio'], i['step'], i['indicator'])
    c = {k: v.cuda() for k, v in i.items()}
    got = m.denoising_net(c['motion'], c['audio'], c['person'], c['style'], c['prev_motion'], c['prev_audio'],
                          c['step'], c['indicator'])
    assert rel_l2(got, want) < X0_TOL
    # per-sequence independence: sequence 0 alone gives the same rows
    got1 = m.denoising_net(c['motion'][:1], c['audio'][:1], c['person'][:1], c['style'][:1], c['prev_motion'][:1],
                           c['prev_audio'][:1], c['step'][:1], c['indicator'][:1])
    assert rel_l2(got1, got[:1]) < 1e-6


@pytest.mark.parametrize('mode', ['incremental', 'independent'])
def test_sampler_teacher_forced_steps_match_golden(built_lib, mode):
    c = SAMP_GOLD
    m, args = make_msmd('cuda', n_diff_steps=c['T'])
    sched = D.cosine_schedule(c['T'])
    i = synth.sampler_inputs(c['N'], c['T'], c['seed'])
    gold = np.load(os.path.join(GOLDEN, 'sampler.npz'))[mode]
    worst = 0.0
    for t in range(c['T'], 0, -1):
        x_t = torch.from_numpy(gold[t]).cuda()
        x_prev, _, _ = m.sample(i['audio_feat'].cuda(), i['shape'].cuda(), i['style'].cuda(), motion_at_T=x_t,
                                indicator=i['indicator'].cuda(), cfg_mode=mode, cfg_scale=list(c['scales']),
                                noise=i['z'].cuda(), t_start=t, n_steps=1)
        a, ab, abp = sched['alphas'][t], sched['alpha_bars'][t], sched['alpha_bars'][t - 1]
        c1 = float((1 - a) * torch.sqrt(abp) / (1 - ab))
        err = rel_l2(x_prev, gold[t - 1])
        worst = max(worst, err / (1e-3 + X0_TOL * c1))
        assert err < 1e-3 + X0_TOL * c1, (t, err, c1)
    print('worst teacher-forced step error / bound:', worst)


def test_sampler_free_running_and_graph_replay(built_lib):
    """All T steps in one call (CUDA-graph replay path) vs the golden trajectory; the free-running bf16
    trajectory drifts from the fp32 one, so the bound is loose - the strict check is the teacher-forced test."""
    c = SAMP_GOLD
    m, args = make_msmd('cuda', n_diff_steps=c['T'])
    i = synth.sampler_inputs(c['N'], c['T'], c['seed'])
    gold = np.load(os.path.join(GOLDEN, 'sampler.npz'))['incremental']
    kw = dict(motion_at_T=i['x_T'].cuda(), indicator=i['indicator'].cuda(), cfg_mode='incremental',
              cfg_scale=list(c['scales']), noise=i['z'].cuda())
    traj, xT, af = m.sample(i['audio_feat'].cuda(), i['shape'].cuda(), i['style'].cuda(), ret_traj=True, **kw)
    assert sorted(traj) == list(range(c['T'] + 1))
    errs = [rel_l2(traj[t], gold[t]) for t in range(c['T'], -1, -1)]
    print('free-running rel-L2 per step:', ['%.1e' % e for e in errs])
    assert errs[0] == 0.0 and max(errs) < 5e-2
    x0, _, _ = m.sample(i['audio_feat'].cuda(), i['shape'].cuda(), i['style'].cuda(), **kw)
    assert torch.equal(x0.cpu(), traj[0].cpu())          # deterministic; graph path == returned trajectory end
    # in-kernel Philox noise path: reproducible from torch's seed, different across seeds
    kw.pop('noise')
    torch.manual_seed(7); a, _, _ = m.sample(i['audio_feat'].cuda(), i['shape'].cuda(), i['style'].cuda(), **kw)
    torch.manual_seed(7); b, _, _ = m.sample(i['audio_feat'].cuda(), i['shape'].cuda(), i['style'].cuda(), **kw)
    torch.manual_seed(8); d, _, _ = m.sample(i['audio_feat'].cuda(), i['shape'].cuda(), i['style'].cuda(), **kw)
    assert torch.equal(a, b) and not torch.equal(a, d) and torch.isfinite(a).all()
