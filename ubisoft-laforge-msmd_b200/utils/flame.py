"""Drop-in for /root/reference/utils/flame.py: FLAME(config).forward(...) -> (vertices, lm2d, lm3d).

Same constructor inputs (FLAME2020 pickle + landmark-embedding npy, flame.py:66-124), same
buffer / parameter names in ``state_dict``, same forward signature
(flame.py:180-181).  The decode itself is csrc/flame*.cu behind the C ABI.
"""
import ctypes as C
import pickle
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .lbs import FlameHandle, vertices2landmarks

FLAMEConfig = SimpleNamespace(
    flame_model_path="/code/models/flame_data/FLAME2020/generic_model.pkl",
    n_shape=100, n_exp=50, n_tex=50, tex_type='BFM',
    tex_path="/code/models/flame_data/FLAME2020/FLAME_albedo_from_BFM.npz",
    flame_lmk_embedding_path='/code/models/flame_data/landmark_embedding.npy')


def _np(a, dtype=np.float32):
    if 'scipy.sparse' in str(type(a)):
        a = a.todense()
    return np.array(a, dtype=dtype)


class FLAME(nn.Module):
    def __init__(self, config, raw=None, lmk_embeddings=None):
        """``config`` as in the reference.  ``raw`` / ``lmk_embeddings`` let callers hand over the
        already un-pickled dicts (synthetic assets in tests/bench) instead of file paths."""
        super().__init__()
        if raw is None:
            with open(config.flame_model_path, 'rb') as f:
                raw = pickle.load(f, encoding='latin1')
        self.dtype = torch.float32
        t = lambda a: torch.tensor(_np(a), dtype=torch.float32)
        self.register_buffer('faces_tensor', torch.tensor(_np(raw['f'], np.int64), dtype=torch.long))
        self.register_buffer('v_template', t(raw['v_template']))
        sd = t(raw['shapedirs'])
        self.register_buffer('shapedirs', torch.cat([sd[:, :, :config.n_shape],
                                                     sd[:, :, 300:300 + config.n_exp]], 2).contiguous())
        pd = _np(raw['posedirs'])
        self.register_buffer('posedirs', torch.tensor(np.reshape(pd, [-1, pd.shape[-1]]).T.copy()))
        self.register_buffer('J_regressor', t(raw['J_regressor']))
        parents = torch.tensor(_np(raw['kintree_table'], np.int64)[0]).long()
        parents[0] = -1
        self.register_buffer('parents', parents)
        self.register_buffer('lbs_weights', t(raw['weights']))
        self.register_parameter('eye_pose', nn.Parameter(torch.zeros(1, 6), requires_grad=False))
        self.register_parameter('eye_pose_mat', nn.Parameter(torch.eye(3).view(1, 9).repeat(1, 2), requires_grad=False))
        self.register_parameter('neck_pose', nn.Parameter(torch.zeros(1, 3), requires_grad=False))
        self.register_parameter('neck_pose_mat', nn.Parameter(torch.eye(3).view(1, 9), requires_grad=False))
        if lmk_embeddings is None and getattr(config, 'flame_lmk_embedding_path', None):
            lmk_embeddings = np.load(config.flame_lmk_embedding_path, allow_pickle=True, encoding='latin1')[()]
        self.has_landmarks = lmk_embeddings is not None
        if self.has_landmarks:
            e = lmk_embeddings
            tt = lambda a, dt: torch.as_tensor(np.asarray(a)).to(dt)
            self.register_buffer('lmk_faces_idx', tt(e['static_lmk_faces_idx'], torch.long))
            self.register_buffer('lmk_bary_coords', tt(e['static_lmk_bary_coords'], torch.float32))
            self.register_buffer('dynamic_lmk_faces_idx', tt(e['dynamic_lmk_faces_idx'], torch.long))
            self.register_buffer('dynamic_lmk_bary_coords', tt(e['dynamic_lmk_bary_coords'], torch.float32))
            self.register_buffer('full_lmk_faces_idx', tt(e['full_lmk_faces_idx'], torch.long))
            self.register_buffer('full_lmk_bary_coords', tt(e['full_lmk_bary_coords'], torch.float32))
        chain, cur = [], 1
        while cur != -1:
            chain.append(cur)
            cur = int(parents[cur])
        self.register_buffer('neck_kin_chain', torch.tensor(chain, dtype=torch.long))
        self._handle = None
        self.impl = 0

    def _flame_handle(self):
        dev = self.shapedirs.device
        if self._handle is None or self._handle.device != dev:
            self._handle = FlameHandle(self.v_template, self.shapedirs, self.posedirs, self.J_regressor,
                                       self.parents, self.lbs_weights, dev)
        return self._handle

    def _contour_rows(self, full_pose, pose2rot):
        """flame.py:126-172: yaw-dependent contour row per frame (one tiny kernel)."""
        B = full_pose.shape[0]
        out = torch.empty(B, dtype=torch.long, device=full_pose.device)
        fp = _lib.as_f32c(full_pose)
        with torch.cuda.device(fp.device):
            _lib.check(_lib.lib().msmd_flame_contour_index(_lib.dev_ptr(fp), int(pose2rot), 5,
                                                           _lib.dev_ptr(self.neck_kin_chain, torch.int64),
                                                           int(self.neck_kin_chain.numel()), B,
                                                           _lib.dev_ptr(out, torch.int64), _lib.stream_ptr()))
        return out

    def seletec_3d68(self, vertices):
        B = vertices.shape[0]
        return vertices2landmarks(vertices, self.faces_tensor, self.full_lmk_faces_idx.repeat(B, 1),
                                  self.full_lmk_bary_coords.repeat(B, 1, 1))

    def forward(self, shape_params=None, expression_params=None, pose_params=None, eye_pose_params=None,
                pose2rot=True, ignore_global_rot=False, return_lm2d=True, return_lm3d=True):
        B = shape_params.shape[0]
        betas = torch.cat([shape_params, expression_params], dim=1)
        if pose2rot:   # flame.py:193-203
            if pose_params is None:
                pose_params = self.eye_pose.expand(B, -1)
            if eye_pose_params is None:
                eye_pose_params = self.eye_pose.expand(B, -1)
            head = pose_params[:, :3] if not ignore_global_rot else torch.zeros_like(pose_params[:, :3])
            full_pose = torch.cat([head, self.neck_pose.expand(B, -1), pose_params[:, 3:], eye_pose_params], dim=1)
        else:          # flame.py:204-211
            if pose_params is None:
                pose_params = self.eye_pose_mat.expand(B, -1)
            if eye_pose_params is None:
                eye_pose_params = self.eye_pose_mat.expand(B, -1)
            head = pose_params[:, :9] if not ignore_global_rot else self.eye_pose_mat.expand(B, -1)[:, :9]
            full_pose = torch.cat([head, self.neck_pose_mat.expand(B, -1), pose_params[:, 9:], eye_pose_params], dim=1)
        vertices, _ = self._flame_handle().decode(betas, full_pose, pose2rot, want_joints=False, impl=self.impl)

        lm2d = lm3d = None
        if return_lm2d:   # flame.py:219-233
            if not self.has_landmarks:
                raise ValueError('FLAME was built without landmark embeddings')
            rows = self._contour_rows(full_pose, pose2rot)
            idx = torch.cat([self.dynamic_lmk_faces_idx[rows], self.lmk_faces_idx.unsqueeze(0).expand(B, -1)], 1)
            bc = torch.cat([self.dynamic_lmk_bary_coords[rows], self.lmk_bary_coords.unsqueeze(0).expand(B, -1, -1)], 1)
            lm2d = vertices2landmarks(vertices, self.faces_tensor, idx, bc)
        if return_lm3d:   # flame.py:237-241
            if not self.has_landmarks:
                raise ValueError('FLAME was built without landmark embeddings')
            lm3d = self.seletec_3d68(vertices)
        return vertices, lm2d, lm3d
