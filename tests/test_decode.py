"""Decode adapter: sampled codes -> FLAME inputs -> vertices.
CPU: oracle/decode.py vs the golden vectors made with the unmodified reference (utils/common.py:140-196 for the 54-d layout;
inference.py:274-275 + rotation_conversions + FLAME.forward for MSMD's 67-d codes).  GPU: the drop-in against both."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import decode as OD, synth
from oracle.make_golden import decode_inputs

TOL = 1e-5


def _gold():
    return np.load(os.path.join(GOLDEN, 'decode.npz'))


def test_oracle_decode_matches_reference_golden():
    g = _gold()
    motion54, shape, stats54, codes67, stats67 = decode_inputs()
    a100 = synth.flame_assets(0, synth.FLAME_V, 100, 50)
    for name, kw in (('global', True), ('noglobal', False)):
        cd = OD.get_coef_dict(motion54, shape, stats54, with_global_pose=kw)
        for k, v in cd.items():
            assert np.array_equal(v.numpy(), g[f'cd_{name}_{k}']), (name, k)
        assert rel_l2(OD.coef_dict_to_vertices(cd, a100), g[f'verts54_{name}']) < 2e-6
    cd = OD.get_coef_dict(motion54, shape, stats54, with_global_pose=True)
    assert rel_l2(OD.coef_dict_to_vertices(cd, a100, ignore_global_rot=True), g['verts54_ignore_global']) < 2e-6
    a300 = synth.flame_assets(0, synth.FLAME_V, 300, 100)
    _, pose = OD.codes_to_flame_inputs(codes67, 100, **stats67)
    assert np.abs(pose[:, :3].numpy() - g['aa67']).max() < 2e-6 and float(pose[:, 3:].abs().max()) == 0.0
    assert rel_l2(OD.decode_vertices(a300, codes67, 300, 100, **stats67), g['verts67']) < 2e-6


def _flame(n_shape, n_exp):
    from types import SimpleNamespace
    from msmd_b200.utils.flame import FLAME
    raw = synth.flame_raw(0, synth.FLAME_V, 400)
    return FLAME(SimpleNamespace(n_shape=n_shape, n_exp=n_exp, flame_lmk_embedding_path=None), raw=raw).cuda()


@pytest.mark.gpu
def test_coef_dict_to_vertices_cuda_matches_reference_golden(built_lib):
    from msmd_b200.utils.common import coef_dict_to_vertices, get_coef_dict
    g = _gold()
    motion54, shape, stats54, _, _ = decode_inputs()
    dev = lambda d: {k: v.cuda() for k, v in d.items()}
    fl = _flame(100, 50)
    for name, kw in (('global', True), ('noglobal', False)):
        cd = get_coef_dict(motion54.cuda(), shape.cuda(), dev(stats54), with_global_pose=kw)
        for k, v in cd.items():
            assert np.array_equal(v.cpu().numpy(), g[f'cd_{name}_{k}']), (name, k)
        for bs in (512, 3):                       # the reference's batching knob does not change values
            v = coef_dict_to_vertices(cd, fl, flame_batch_size=bs)
            assert v.shape == (2, 4, synth.FLAME_V, 3) and rel_l2(v, g[f'verts54_{name}']) < TOL, (name, bs)
    cd = get_coef_dict(motion54.cuda(), shape.cuda(), dev(stats54), with_global_pose=True)
    assert rel_l2(coef_dict_to_vertices(cd, fl, ignore_global_rot=True), g['verts54_ignore_global']) < TOL
    with pytest.raises(ValueError, match='Unknown rot'):
        coef_dict_to_vertices(cd, fl, rot_repr='euler')
    with pytest.raises(ValueError, match='Unknown rotation'):
        get_coef_dict(motion54.cuda(), rot_repr='6d')


@pytest.mark.gpu
def test_decode_vertices_cuda_matches_reference_golden_and_oracle(built_lib):
    from msmd_b200.decode import codes_to_flame_inputs, decode_vertices
    g = _gold()
    _, _, _, codes67, stats67 = decode_inputs()
    st = {k: v.cuda() for k, v in stats67.items()}
    fl = _flame(300, 100)
    expression, pose = codes_to_flame_inputs(codes67.cuda(), 100, **st)
    assert expression.shape == (8, 100) and float(expression[:, 64:].abs().max()) == 0.0
    assert np.abs(pose[:, :3].cpu().numpy() - g['aa67']).max() < 3e-5      # matrix_to_axis_angle is ill-conditioned near pi
    v = decode_vertices(fl, codes67.cuda(), n_exp=100, **st)
    err = rel_l2(v, g['verts67'])
    print('decode_vertices (67-d codes -> vertices) rel-L2 vs reference golden:', err)
    assert v.shape == (2, 4, synth.FLAME_V, 3) and err < TOL
    # a larger batch without statistics (what bench.py times) against the oracle
    codes = torch.randn(3, 50, 67, generator=torch.Generator().manual_seed(5))
    want = OD.decode_vertices(synth.flame_assets(0, synth.FLAME_V, 300, 100), codes, 300, 100)
    assert rel_l2(decode_vertices(fl, codes.cuda(), n_exp=100), want) < TOL
