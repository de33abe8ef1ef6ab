"""CPU restatement (torch-CPU, functional) of the decode glue that follows the sampler.  TEST INFRASTRUCTURE ONLY:
nothing in the product package or bench.py's GPU arm imports it.

Two layouts:
  * the reference's 54-d DiffPoseTalk coefficient layout: ``get_coef_dict`` / ``coef_dict_to_vertices``
    (/root/reference/utils/common.py:140-196);
  * MSMD's own 67-d codes (64 expression + head rotation as Euler 'YXZ' degrees, dataset_processing/Step2*.py:556-566,
    :657-659): de-normalise (inference.py:274-275) -> euler_angles_to_matrix -> matrix_to_axis_angle
    (utils/rotation_conversions.py:151, :434) -> FLAME.forward (utils/flame.py:180).  The reference leaves this last
    stage to the user (inference.py:281-283); every piece of it is a reference function.
Pinned by tests/golden/decode.npz (oracle/make_golden.py gen_decode, produced with the unmodified reference).
"""
import math

import torch

from . import flame_lbs, rotations


def get_coef_dict(motion_coef, shape_coef=None, denorm_stats=None, with_global_pose=False):
    """common.py:140-173 (rot_repr == 'aa')."""
    coef = {'exp': motion_coef[..., :50]}
    if with_global_pose:
        pose = motion_coef[..., 50:]
    else:
        pose = torch.cat([torch.zeros_like(motion_coef[..., :3]), motion_coef[..., -1:]], -1)      # :148-149
    coef['pose'] = torch.cat([pose, torch.zeros_like(motion_coef[..., :2])], -1)                    # :151
    if shape_coef is not None:
        if motion_coef.ndim == 3:                                                                   # :156-160
            if shape_coef.ndim == 2:
                shape_coef = shape_coef.unsqueeze(1)
            if shape_coef.shape[1] == 1:
                shape_coef = shape_coef.expand(-1, motion_coef.shape[1], -1)
        coef['shape'] = shape_coef
    if denorm_stats is not None:                                                                    # :164-165
        coef = {k: coef[k] * denorm_stats[f'{k}_std'] + denorm_stats[f'{k}_mean'] for k in coef}
    if not with_global_pose:                                                                        # :167-169
        coef['pose'] = coef['pose'].clone()
        coef['pose'][..., :3] = 0
    return coef


def coef_dict_to_vertices(coef, assets, ignore_global_rot=False):
    """common.py:176-196: FLAME.forward on the flattened frames (the 512-frame batching does not change values)."""
    shape = coef['exp'].shape[:-1]
    flat = {k: v.reshape(-1, v.shape[-1]) for k, v in coef.items()}
    v = flame_lbs.flame_forward(assets, flat['shape'], flat['exp'], flat['pose'], None, ignore_global_rot=ignore_global_rot)
    return v.view(*shape, -1, 3)


def codes_to_flame_inputs(codes, n_exp, exp_mean=None, exp_std=None, rot_mean=None, rot_std=None):
    """codes [N,F,67] -> (expression [N*F,n_exp], pose [N*F,6] = global axis-angle | jaw zeros)."""
    N, Fr, D = codes.shape
    exp, rot = codes[..., :D - 3], codes[..., D - 3:]
    if exp_std is not None:
        exp = exp * exp_std + exp_mean                                                              # inference.py:274
    if rot_std is not None:
        rot = rot * rot_std + rot_mean                                                              # inference.py:275
    expression = torch.zeros(N * Fr, n_exp)
    k = min(n_exp, D - 3)
    expression[:, :k] = exp.reshape(N * Fr, -1)[:, :k]
    rad = rot.reshape(N * Fr, 3) * (math.pi / 180.0)                                                # Step2:563 (degrees)
    aa = torch.as_tensor(rotations.matrix_to_axis_angle(rotations.euler_angles_to_matrix(rad.numpy(), 'YXZ')),
                         dtype=torch.float32)
    return expression, torch.cat([aa, torch.zeros_like(aa)], 1)


def decode_vertices(assets, codes, n_shape, n_exp, **denorm):
    N, Fr, _ = codes.shape
    expression, pose = codes_to_flame_inputs(codes, n_exp, **denorm)
    v = flame_lbs.flame_forward(assets, torch.zeros(N * Fr, n_shape), expression, pose, None)
    return v.view(N, Fr, -1, 3)
