"""Shared builders for the tests (seeded weights + modules)."""
import torch
import torch.nn as nn

from oracle import synth
from oracle.ref_shims import pinned_args


def make_msmd(device='cpu', weight_seed=1234, precision='bf16', **over):
    """The drop-in MSMD with deterministic weights (same fill as oracle.make_golden.ref_msmd).  ``precision`` defaults to
    the explicit pure-bf16 engine here (most tests pin one arithmetic); pass None for the package default ('hybrid')."""
    from msmd_b200 import model as M
    args = pinned_args(**over)
    m = M.MSMD(args, 'cpu', True, use_head_alpha=False, regularize_alpha="None", audio_encoder=nn.Identity())
    fill = synth.fill_state_dict(synth.param_spec(m), weight_seed)
    missing, unexpected = m.load_state_dict(fill, strict=False)
    assert not unexpected
    if precision is not None:
        m.precision = m.denoising_net.precision = precision
    return m.to(device).eval(), args


def cpu_state_dict(module):
    return {k: v.detach().cpu() for k, v in module.state_dict().items()}
