"""Decode glue of /root/reference/utils/common.py (:140-196) for the 54-d DiffPoseTalk coefficient layout
(50 expression + global pose (3) + jaw (1)): tensor slicing on the host, FLAME through the CUDA decode."""
from functools import reduce

import torch


def get_coef_dict(motion_coef, shape_coef=None, denorm_stats=None, with_global_pose=False, rot_repr='aa'):
    """common.py:140-173."""
    if rot_repr != 'aa':
        raise ValueError(f'Unknown rotation representation {rot_repr}!')
    coef = {'exp': motion_coef[..., :50]}
    pose = motion_coef[..., 50:] if with_global_pose else torch.cat(
        [torch.zeros_like(motion_coef[..., :3]), motion_coef[..., -1:]], dim=-1)
    coef['pose'] = torch.cat([pose, torch.zeros_like(motion_coef[..., :2])], dim=-1)   # jaw y/z rotations back as 0
    if shape_coef is not None:
        if motion_coef.ndim == 3:
            if shape_coef.ndim == 2:
                shape_coef = shape_coef.unsqueeze(1)
            if shape_coef.shape[1] == 1:
                shape_coef = shape_coef.expand(-1, motion_coef.shape[1], -1)
        coef['shape'] = shape_coef
    if denorm_stats is not None:
        coef = {k: coef[k] * denorm_stats[f'{k}_std'] + denorm_stats[f'{k}_mean'] for k in coef}
    if not with_global_pose:
        coef['pose'][..., :3] = 0
    return coef


def coef_dict_to_vertices(coef_dict, flame, rot_repr='aa', ignore_global_rot=False, flame_batch_size=512):
    """common.py:176-196.  The reference loops over 512-frame FLAME batches to bound its ~10 GB of intermediates;
    the fused decode has none, so the batch size only bounds the output tensor of one call."""
    if rot_repr != 'aa':
        raise ValueError(f'Unknown rot_repr: {rot_repr}')
    shape = coef_dict['exp'].shape[:-1]
    flat = {k: v.reshape(-1, v.shape[-1]) for k, v in coef_dict.items()}
    n = reduce(lambda x, y: x * y, shape, 1)
    step = max(int(flame_batch_size), 1) * 64
    out = []
    for i in range(0, n, step):
        v, _, _ = flame(flat['shape'][i:i + step], flat['exp'][i:i + step], flat['pose'][i:i + step], pose2rot=True,
                        ignore_global_rot=ignore_global_rot, return_lm2d=False, return_lm3d=False)
        out.append(v)
    return torch.cat(out, dim=0).view(*shape, -1, 3)
