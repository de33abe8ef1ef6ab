"""Drop-in for /root/reference/utils/rotation_conversions.py (same names, shapes, exceptions).

Each function validates like the reference, flattens the leading dims and runs ONE
fused sm_100a kernel through the C ABI (msmd_rot_convert / msmd_quat_binary) instead of
the reference's 10-30 ATen ops and boolean-mask indexing (which synchronises the device,
rotation_conversions.py:92-94, :466-475).  CUDA fp32 tensors only.
"""
from typing import Optional

import torch

from .. import _lib

_K = dict(QUAT_TO_MATRIX=0, MATRIX_TO_QUAT=1, EULER_TO_MATRIX=2, MATRIX_TO_EULER=3, AA_TO_QUAT=4,
          QUAT_TO_AA=5, AA_TO_MATRIX=6, MATRIX_TO_AA=7, R6D_TO_MATRIX=8, MATRIX_TO_6D=9, AA_TO_6D=10,
          STANDARDIZE_QUAT=11, QUAT_INVERT=12, EULER_TO_AA=13, RODRIGUES=14)
_AX = {'X': 0, 'Y': 1, 'Z': 2}


def _convention_code(convention: str) -> int:
    # rotation_conversions.py:163-171 / :229-237
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError(f"Invalid convention {convention}.")
    for letter in convention:
        if letter not in ("X", "Y", "Z"):
            raise ValueError(f"Invalid letter {letter} in convention string.")
    return _AX[convention[0]] * 9 + _AX[convention[1]] * 3 + _AX[convention[2]]


def _run(kind, x, in_w, out_shape_tail, conv=0):
    lead = x.shape[:-1] if in_w != 9 else x.shape[:-2]
    xin = _lib.as_f32c(x).reshape(-1, in_w)
    n = xin.shape[0]
    out_w = 1
    for d in out_shape_tail:
        out_w *= d
    out = torch.empty((n, out_w), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().msmd_rot_convert(kind, _lib.dev_ptr(xin), _lib.dev_ptr(out), n, conv, _lib.stream_ptr()))
    return out.reshape(tuple(lead) + tuple(out_shape_tail))


def _check_matrix(matrix):
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix  shape f{matrix.shape}.")


def quaternion_to_matrix(quaternions):
    return _run(_K['QUAT_TO_MATRIX'], quaternions, 4, (3, 3))


def matrix_to_quaternion(matrix):
    _check_matrix(matrix)
    return _run(_K['MATRIX_TO_QUAT'], matrix, 9, (4,))


def euler_angles_to_matrix(euler_angles, convention: str):
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    return _run(_K['EULER_TO_MATRIX'], euler_angles, 3, (3, 3), _convention_code(convention))


def matrix_to_euler_angles(matrix, convention: str):
    code = _convention_code(convention)
    _check_matrix(matrix)
    return _run(_K['MATRIX_TO_EULER'], matrix, 9, (3,), code)


def axis_angle_to_quaternion(axis_angle):
    return _run(_K['AA_TO_QUAT'], axis_angle, 3, (4,))


def quaternion_to_axis_angle(quaternions):
    return _run(_K['QUAT_TO_AA'], quaternions, 4, (3,))


def axis_angle_to_matrix(axis_angle):
    return _run(_K['AA_TO_MATRIX'], axis_angle, 3, (3, 3))


def matrix_to_axis_angle(matrix):
    _check_matrix(matrix)
    return _run(_K['MATRIX_TO_AA'], matrix, 9, (3,))


def rotation_6d_to_matrix(d6: torch.Tensor) -> torch.Tensor:
    return _run(_K['R6D_TO_MATRIX'], d6, 6, (3, 3))


def matrix_to_rotation_6d(matrix: torch.Tensor) -> torch.Tensor:
    return _run(_K['MATRIX_TO_6D'], matrix, 9, (6,))


def axis_angle_to_rotation_6d(axis_angle):
    return _run(_K['AA_TO_6D'], axis_angle, 3, (6,))


def standardize_quaternion(quaternions):
    return _run(_K['STANDARDIZE_QUAT'], quaternions, 4, (4,))


def quaternion_invert(quaternion):
    return _run(_K['QUAT_INVERT'], quaternion, 4, (4,))


def euler_angles_to_axis_angle(euler_angles, convention: str):
    """Fused euler -> matrix -> axis-angle (== matrix_to_axis_angle(euler_angles_to_matrix(.)));
    the head-pose step of the decode adapter."""
    if euler_angles.dim() == 0 or euler_angles.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    return _run(_K['EULER_TO_AA'], euler_angles, 3, (3,), _convention_code(convention))


def _binary(op, a, b, bw, ow):
    lead = torch.broadcast_shapes(a.shape[:-1], b.shape[:-1])
    A = _lib.as_f32c(a.expand(lead + (4,))).reshape(-1, 4)
    Bm = _lib.as_f32c(b.expand(lead + (bw,))).reshape(-1, bw)
    out = torch.empty((A.shape[0], ow), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().msmd_quat_binary(op, _lib.dev_ptr(A), _lib.dev_ptr(Bm), _lib.dev_ptr(out),
                                               A.shape[0], _lib.stream_ptr()))
    return out.reshape(lead + (ow,))


def quaternion_raw_multiply(a, b):
    return _binary(0, a, b, 4, 4)


def quaternion_multiply(a, b):
    return _binary(1, a, b, 4, 4)


def quaternion_apply(quaternion, point):
    if point.size(-1) != 3:
        raise ValueError(f"Points are not in 3D, f{point.shape}.")
    return _binary(2, quaternion, point, 3, 3)


def random_quaternions(n: int, dtype: Optional[torch.dtype] = None, device=None, requires_grad=False):
    """rotation_conversions.py:260-283 (RNG stays torch's; normalisation sign = sign of w)."""
    o = torch.randn((n, 4), dtype=dtype, device=device)
    s = (o * o).sum(1).sqrt()
    s = torch.where((s < 0) != (o[:, 0] < 0), -s, s)
    return o / s[:, None]


def random_rotations(n: int, dtype: Optional[torch.dtype] = None, device=None, requires_grad=False):
    return quaternion_to_matrix(random_quaternions(n, dtype=dtype, device=device))


def random_rotation(dtype: Optional[torch.dtype] = None, device=None, requires_grad=False):
    return random_rotations(1, dtype, device)[0]
