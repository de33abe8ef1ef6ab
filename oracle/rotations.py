"""Oracle: rotation conversions (numpy float32 restatement).

Restates /root/reference/utils/rotation_conversions.py (cited per function).
All functions broadcast over leading dims and compute in the input dtype
(float32 in the tests).  Quirks kept on purpose: the
sqrt/copysign ``matrix_to_quaternion`` and the two-branch small-angle series.
"""
import numpy as np

_AX = {"X": 0, "Y": 1, "Z": 2}


def _f(x):
    return np.asarray(x)


def check_convention(convention):
    """Same ValueErrors as rotation_conversions.py:163-171 / :229-237."""
    if len(convention) != 3:
        raise ValueError("Convention must have 3 letters.")
    if convention[1] in (convention[0], convention[2]):
        raise ValueError(f"Invalid convention {convention}.")
    for letter in convention:
        if letter not in ("X", "Y", "Z"):
            raise ValueError(f"Invalid letter {letter} in convention string.")


def quaternion_to_matrix(q):
    """rotation_conversions.py:38-67 (two_s = 2/|q|^2, no normalisation of q)."""
    q = _f(q)
    r, i, j, k = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    two_s = q.dtype.type(2.0) / (q * q).sum(-1)
    one = q.dtype.type(1.0)
    o = np.stack([
        one - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), one - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), one - two_s * (i * i + j * j),
    ], -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _copysign(a, b):
    """rotation_conversions.py:70-86: flips when (a<0) != (b<0); b==0 counts as 'not negative'."""
    return np.where((a < 0) != (b < 0), -a, a)


def _sqrt_pos(x):
    """rotation_conversions.py:89-97: sqrt(max(0,x)), exactly 0 where x<=0."""
    out = np.zeros_like(x)
    m = x > 0
    out[m] = np.sqrt(x[m])
    return out


def matrix_to_quaternion(m):
    """rotation_conversions.py:100-120 (old, lossy-near-180deg formula)."""
    m = _f(m)
    if m.shape[-1] != 3 or m.shape[-2] != 3:
        raise ValueError(f"Invalid rotation matrix  shape f{m.shape}.")
    h = m.dtype.type(0.5)
    one = m.dtype.type(1.0)
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    w = h * _sqrt_pos(one + m00 + m11 + m22)
    x = h * _sqrt_pos(one + m00 - m11 - m22)
    y = h * _sqrt_pos(one - m00 + m11 - m22)
    z = h * _sqrt_pos(one - m00 - m11 + m22)
    x = _copysign(x, m[..., 2, 1] - m[..., 1, 2])
    y = _copysign(y, m[..., 0, 2] - m[..., 2, 0])
    z = _copysign(z, m[..., 1, 0] - m[..., 0, 1])
    return np.stack([w, x, y, z], -1)


def _axis_rot(axis, ang):
    """rotation_conversions.py:123-148."""
    c, s = np.cos(ang), np.sin(ang)
    o, z = np.ones_like(ang), np.zeros_like(ang)
    if axis == "X":
        flat = (o, z, z, z, c, -s, z, s, c)
    elif axis == "Y":
        flat = (c, z, s, z, o, z, -s, z, c)
    else:
        flat = (c, -s, z, s, c, z, z, z, o)
    return np.stack(flat, -1).reshape(ang.shape + (3, 3))


def euler_angles_to_matrix(e, convention):
    """rotation_conversions.py:151-173: R = R_c0(e0) @ R_c1(e1) @ R_c2(e2)."""
    e = _f(e)
    if e.ndim == 0 or e.shape[-1] != 3:
        raise ValueError("Invalid input euler angles.")
    check_convention(convention)
    m = [_axis_rot(c, e[..., n]) for n, c in enumerate(convention)]
    return np.matmul(np.matmul(m[0], m[1]), m[2])


def _angle_from_tan(axis, other_axis, data, horizontal, tait_bryan):
    """rotation_conversions.py:176-207."""
    i1, i2 = {"X": (2, 1), "Y": (0, 2), "Z": (1, 0)}[axis]
    if horizontal:
        i2, i1 = i1, i2
    even = (axis + other_axis) in ["XY", "YZ", "ZX"]
    if horizontal == even:
        return np.arctan2(data[..., i1], data[..., i2])
    if tait_bryan:
        return np.arctan2(-data[..., i2], data[..., i1])
    return np.arctan2(data[..., i2], -data[..., i1])


def matrix_to_euler_angles(m, convention):
    """rotation_conversions.py:219-257."""
    m = _f(m)
    check_convention(convention)
    if m.shape[-1] != 3 or m.shape[-2] != 3:
        raise ValueError(f"Invalid rotation matrix  shape f{m.shape}.")
    i0, i2 = _AX[convention[0]], _AX[convention[2]]
    tb = i0 != i2
    if tb:
        sign = m.dtype.type(-1.0 if (i0 - i2) in (-1, 2) else 1.0)
        central = np.arcsin(m[..., i0, i2] * sign)
    else:
        central = np.arccos(m[..., i0, i0])
    o = (_angle_from_tan(convention[0], convention[1], m[..., i2], False, tb),
         central,
         _angle_from_tan(convention[2], convention[1], m[..., i0, :], True, tb))
    return np.stack(o, -1)


def standardize_quaternion(q):
    """rotation_conversions.py:326-338."""
    q = _f(q)
    return np.where(q[..., 0:1] < 0, -q, q)


def quaternion_raw_multiply(a, b):
    """rotation_conversions.py:341-359 (Hamilton product, real part first)."""
    a, b = np.broadcast_arrays(_f(a), _f(b))
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw], -1)


def quaternion_multiply(a, b):
    """rotation_conversions.py:362-377."""
    return standardize_quaternion(quaternion_raw_multiply(a, b))


def quaternion_invert(q):
    """rotation_conversions.py:380-393."""
    q = _f(q)
    return q * np.array([1, -1, -1, -1], dtype=q.dtype)


def quaternion_apply(q, p):
    """rotation_conversions.py:396-415: (q * (0,p)) * conj(q), vector part."""
    q, p = _f(q), _f(p)
    if p.shape[-1] != 3:
        raise ValueError(f"Points are not in 3D, f{p.shape}.")
    lead = np.broadcast_shapes(q.shape[:-1], p.shape[:-1])
    pq = np.concatenate([np.zeros(p.shape[:-1] + (1,), p.dtype), p], -1)
    out = quaternion_raw_multiply(quaternion_raw_multiply(np.broadcast_to(q, lead + (4,)),
                                                          np.broadcast_to(pq, lead + (4,))),
                                  quaternion_invert(q))
    return out[..., 1:]


def _half_sinc(angles, half):
    """sin(a/2)/a with the a<1e-6 series branch (rotation_conversions.py:462-475, :498-509)."""
    small = np.abs(angles) < 1e-6
    safe = np.where(small, np.ones_like(angles), angles)
    big = np.sin(half) / safe
    ser = angles.dtype.type(0.5) - (angles * angles) / angles.dtype.type(48)
    return np.where(small, ser, big)


def axis_angle_to_quaternion(aa):
    """rotation_conversions.py:450-478."""
    aa = _f(aa)
    ang = np.sqrt((aa * aa).sum(-1, keepdims=True))
    half = aa.dtype.type(0.5) * ang
    return np.concatenate([np.cos(half), aa * _half_sinc(ang, half)], -1)


def quaternion_to_axis_angle(q):
    """rotation_conversions.py:481-510."""
    q = _f(q)
    v = q[..., 1:]
    n = np.sqrt((v * v).sum(-1, keepdims=True))
    half = np.arctan2(n, q[..., :1])
    ang = q.dtype.type(2) * half
    return v / _half_sinc(ang, half)


def axis_angle_to_matrix(aa):
    """rotation_conversions.py:418-431."""
    return quaternion_to_matrix(axis_angle_to_quaternion(aa))


def matrix_to_axis_angle(m):
    """rotation_conversions.py:434-447."""
    return quaternion_to_axis_angle(matrix_to_quaternion(m))


def _normalize(v, eps=1e-12):
    n = np.sqrt((v * v).sum(-1, keepdims=True))
    return v / np.maximum(n, v.dtype.type(eps))


def rotation_6d_to_matrix(d6):
    """rotation_conversions.py:513-535 (Gram-Schmidt, F.normalize eps=1e-12)."""
    d6 = _f(d6)
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = _normalize(a1)
    b2 = _normalize(a2 - (b1 * a2).sum(-1, keepdims=True) * b1)
    b3 = np.cross(b1, b2, axis=-1)
    return np.stack([b1, b2, b3], -2)


def matrix_to_rotation_6d(m):
    """rotation_conversions.py:538-552: first two rows, flattened."""
    m = _f(m)
    return m[..., :2, :].reshape(m.shape[:-2] + (6,)).copy()


def axis_angle_to_rotation_6d(aa):
    """rotation_conversions.py:555-569."""
    return matrix_to_rotation_6d(axis_angle_to_matrix(aa))
