"""Drop-in for /root/reference/utils/lbs.py (hot-path functions only).

``lbs`` keeps the reference signature (lbs.py:141-142) and return value (verts, joints) but
runs as two kernels behind the C ABI: a per-frame pose/joint-chain kernel and one fused
blendshape-GEMM + skinning kernel (csrc/flame.cu, csrc/flame_tc.cu).  blend_shapes,
vertices2joints, batch_rigid_transform and transform_mat (lbs.py:226-371) have no standalone
equivalent: they are stages of that fused kernel.
"""
import ctypes as C

import torch

from .. import _lib

_K_RODRIGUES = 14


class FlameHandle:
    """Owns a packed msmd_flame (static bases on one device)."""

    def __init__(self, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise _lib.MsmdError('msmd_b200 FLAME decode needs a CUDA device (no CPU path)')
        self.device = device
        self.V = int(v_template.shape[-2])
        self.NB = int(shapedirs.shape[-1])
        self.NJ = int(J_regressor.shape[0])
        f = lambda t: t.detach().to('cpu', torch.float32).contiguous()
        vt, sd, pd, jr, lw = f(v_template), f(shapedirs), f(posedirs), f(J_regressor), f(lbs_weights)
        par = torch.as_tensor(parents).detach().to('cpu', torch.int64).contiguous()
        if tuple(sd.shape) != (self.V, 3, self.NB) or tuple(pd.shape) != ((self.NJ - 1) * 9, self.V * 3):
            raise ValueError(f'bad FLAME asset shapes: shapedirs {tuple(sd.shape)}, posedirs {tuple(pd.shape)}')
        h = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().msmd_flame_create(vt.data_ptr(), sd.data_ptr(), pd.data_ptr(), jr.data_ptr(),
                                                par.data_ptr(), lw.data_ptr(), self.V, self.NB, self.NJ,
                                                idx, C.byref(h)))
        self._h = h

    def decode(self, betas, pose, pose2rot=True, want_joints=True, impl=0):
        B = max(betas.shape[0], pose.shape[0])
        betas = _lib.as_f32c(betas.expand(B, -1))
        pose = _lib.as_f32c(pose.reshape(pose.shape[0], self.NJ * (3 if pose2rot else 9)).expand(B, -1))
        if betas.shape[1] != self.NB:
            raise ValueError(f'betas has {betas.shape[1]} coefficients, model has {self.NB}')
        if pose.shape[1] != self.NJ * (3 if pose2rot else 9):
            raise ValueError(f'pose has {pose.shape[1]} values for {self.NJ} joints (pose2rot={pose2rot})')
        verts = torch.empty((B, self.V, 3), dtype=torch.float32, device=betas.device)
        joints = torch.empty((B, self.NJ, 3), dtype=torch.float32, device=betas.device) if want_joints else None
        with torch.cuda.device(betas.device):
            _lib.check(_lib.lib().msmd_flame_decode(self._h, _lib.dev_ptr(betas), _lib.dev_ptr(pose), int(pose2rot),
                                                    B, _lib.dev_ptr(verts), _lib.dev_ptr(joints), impl,
                                                    _lib.stream_ptr()))
        return verts, joints

    def __del__(self):
        try:
            if self._h:
                _lib.lib().msmd_flame_destroy(self._h)
                self._h = None
        except Exception:
            pass


_handles = {}


def _cached_handle(v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    vt = v_template[0] if v_template.dim() == 3 else v_template
    key = (vt.data_ptr(), shapedirs.data_ptr(), posedirs.data_ptr(), J_regressor.data_ptr(), lbs_weights.data_ptr(),
           tuple(shapedirs.shape), str(shapedirs.device), vt._version, shapedirs._version)
    h = _handles.get(key)
    if h is None:
        if len(_handles) > 8:
            _handles.clear()
        h = FlameHandle(vt, shapedirs, posedirs, J_regressor, parents, lbs_weights, shapedirs.device)
        _handles[key] = h
    return h


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot=True,
        dtype=torch.float32):
    """lbs.py:141-223.  Static assets are packed once per distinct set of tensors (cached)."""
    h = _cached_handle(v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights)
    return h.decode(betas, pose, pose2rot, True)


def batch_rodrigues(rot_vecs, epsilon=1e-8, dtype=torch.float32):
    """lbs.py:270-301 (angle = ||r + 1e-8||; `epsilon` unused there too)."""
    x = _lib.as_f32c(rot_vecs).reshape(-1, 3)
    out = torch.empty((x.shape[0], 3, 3), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().msmd_rot_convert(_K_RODRIGUES, _lib.dev_ptr(x), _lib.dev_ptr(out), x.shape[0], 0,
                                               _lib.stream_ptr()))
    return out


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary_coords):
    """lbs.py:102-138: vertices [B,V,3], faces [F,3] long, lmk_faces_idx [B,L] long, bary [B,L,3]."""
    B, V = vertices.shape[:2]
    L = lmk_faces_idx.shape[1]
    v = _lib.as_f32c(vertices)
    f = faces.to(torch.int64).contiguous()
    idx = lmk_faces_idx.to(torch.int64).contiguous()
    bc = _lib.as_f32c(lmk_bary_coords)
    out = torch.empty((B, L, 3), dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().msmd_vertices2landmarks(_lib.dev_ptr(v), _lib.dev_ptr(f, torch.int64),
                                                      _lib.dev_ptr(idx, torch.int64), _lib.dev_ptr(bc), B, V, L,
                                                      _lib.dev_ptr(out), _lib.stream_ptr()))
    return out
