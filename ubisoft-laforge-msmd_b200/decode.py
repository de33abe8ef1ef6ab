"""Decode adapter: sampled motion codes -> FLAME inputs -> vertices.

The reference stops at the codes and leaves the mesh decode to the user (inference.py:281-283).  The
last three motion dims are the head rotation as Euler 'YXZ' angles in DEGREES
(dataset_processing/Step2_*.py:556-566, :657-659); the first 64 are expression codes.  De-normalisation
follows inference.py:274-275 (x * std + mean) when statistics are given.
"""
import math

import torch

from .utils.rotation_conversions import euler_angles_to_axis_angle


def codes_to_flame_inputs(codes, n_exp, exp_mean=None, exp_std=None, rot_mean=None, rot_std=None):
    """codes [N, F, 67] -> (expression [N*F, n_exp], pose [N*F, 6]) for FLAME.forward."""
    N, Fr, D = codes.shape
    exp = codes[..., :D - 3]
    rot = codes[..., D - 3:]
    if exp_std is not None:
        exp = exp * exp_std + exp_mean
    if rot_std is not None:
        rot = rot * rot_std + rot_mean
    expression = torch.zeros((N * Fr, n_exp), device=codes.device, dtype=torch.float32)
    k = min(n_exp, D - 3)
    expression[:, :k] = exp.reshape(N * Fr, -1)[:, :k]
    aa = euler_angles_to_axis_angle((rot.reshape(N * Fr, 3) * (math.pi / 180.0)).contiguous(), 'YXZ')
    pose = torch.cat([aa, torch.zeros_like(aa)], dim=1)     # global rotation | jaw (not modelled by the 67-d codes)
    return expression, pose


@torch.no_grad()
def decode_vertices(flame, codes, shape_params=None, **denorm):
    """codes [N, F, 67] -> vertices [N, F, V, 3] through FLAME.forward (utils/flame.py:180)."""
    N, Fr, _ = codes.shape
    n_shape = flame.shapedirs.shape[-1] - 0
    n_exp = denorm.pop('n_exp')
    expression, pose = codes_to_flame_inputs(codes, n_exp, **denorm)
    if shape_params is None:
        shape_params = torch.zeros((N * Fr, n_shape - n_exp), device=codes.device)
    verts, _, _ = flame(shape_params, expression, pose, None, return_lm2d=False, return_lm3d=False)
    return verts.view(N, Fr, -1, 3)
