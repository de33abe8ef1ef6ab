"""Oracle: FLAME decode = blendshapes + pose correctives + joint chain + skinning.

torch-CPU restatement of /root/reference/utils/lbs.py and
/root/reference/utils/flame.py (FLAME.forward).  Written as explicit formulas
(no 4x4 homogeneous bookkeeping) but mathematically identical; order of the
fp32 operations that matter (Rodrigues quirk, chain composition, the final
affine apply) follows the reference lines cited.
"""
import torch


def rodrigues(rot_vecs):
    """lbs.py:270-301.  Quirk kept: angle = ||r + 1e-8|| (lbs.py:285), axis = r/angle.

    rot_vecs [n,3] -> R [n,3,3] = I + sin(a) K + (1-cos(a)) K@K
    """
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)           # [n,1]
    d = rot_vecs / angle
    s = torch.sin(angle)[:, :, None]
    c = torch.cos(angle)[:, :, None]
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    o = torch.zeros_like(x)
    K = torch.stack([o, -z, y, z, o, -x, -y, x, o], 1).view(-1, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def joint_chain(R, J, parents):
    """lbs.py:317-371 (batch_rigid_transform).

    R [B,nj,3,3], J [B,nj,3] rest joints, parents list (parents[0] = -1).
    Returns posed joints [B,nj,3] and the skinning affines (Ag [B,nj,3,3],
    tg [B,nj,3]) with  x_world = Ag x_rest + tg,  i.e. rows 0..2 of the
    reference's ``rel_transforms`` (lbs.py:366-369: t - A@j).
    """
    nj = R.shape[1]
    rel = J.clone()
    for i in range(1, nj):
        rel[:, i] = J[:, i] - J[:, parents[i]]                      # lbs.py:343-344
    G = [R[:, 0]]
    t = [rel[:, 0]]
    for i in range(1, nj):                                           # lbs.py:354-359
        p = parents[i]
        G.append(torch.bmm(G[p], R[:, i]))
        t.append(torch.bmm(G[p], rel[:, i, :, None])[:, :, 0] + t[p])
    G = torch.stack(G, 1)
    t = torch.stack(t, 1)
    tg = t - torch.einsum('bjik,bjk->bji', G, J)                     # lbs.py:366-369
    return t, G, tg


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights,
        pose2rot=True):
    """lbs.py:141-223.

    betas [B,NB]; pose [B,nj*3] axis-angle (pose2rot) or [B,nj*9] matrices;
    v_template [V,3] (or [B,V,3]); shapedirs [V,3,NB]; posedirs [(nj-1)*9, V*3];
    J_regressor [nj,V]; lbs_weights [V,nj].  Returns verts [B,V,3], joints [B,nj,3].
    """
    B = max(betas.shape[0], pose.shape[0])
    nj = J_regressor.shape[0]
    parents = [int(p) for p in parents]
    # lbs.py:185,246-267  v_shaped = template + sum_l betas[b,l] * shapedirs[v,k,l]
    v_shaped = v_template + torch.einsum('bl,mkl->bmk', betas, shapedirs)
    # lbs.py:189,226-243  J = J_regressor @ v_shaped
    J = torch.einsum('bik,ji->bjk', v_shaped, J_regressor)
    # lbs.py:194-198
    if pose2rot:
        R = rodrigues(pose.reshape(-1, 3)).view(B, nj, 3, 3)
    else:
        R = pose.reshape(B, nj, 3, 3)
    # lbs.py:200-204  pose correctives
    eye = torch.eye(3, dtype=betas.dtype)
    pf = (R[:, 1:] - eye).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pf, posedirs).view(B, -1, 3)
    # lbs.py:206
    Jt, Ag, tg = joint_chain(R, J, parents)
    # lbs.py:210-221  per-vertex blended affine, applied to [v_posed;1]
    Tm = torch.einsum('vj,bjik->bvik', lbs_weights, Ag)              # [B,V,3,3]
    tt = torch.einsum('vj,bji->bvi', lbs_weights, tg)                # [B,V,3]
    verts = torch.einsum('bvik,bvk->bvi', Tm, v_posed) + tt
    return verts, Jt


def flame_full_pose(pose_params, eye_pose_params, B, dtype=torch.float32, ignore_global_rot=False):
    """flame.py:193-203: [global(3) | neck(3)=0 | jaw(3)=pose[:,3:6] | eyes(6)].

    neck pose is a zero module parameter (flame.py:99-101); missing pose/eye
    default to zeros (flame.py:197-200).
    """
    if pose_params is None:
        pose_params = torch.zeros(B, 6, dtype=dtype)
    if eye_pose_params is None:
        eye_pose_params = torch.zeros(B, 6, dtype=dtype)
    head = torch.zeros_like(pose_params[:, :3]) if ignore_global_rot else pose_params[:, :3]
    neck = torch.zeros(B, 3, dtype=dtype)
    return torch.cat([head, neck, pose_params[:, 3:], eye_pose_params], 1)


def flame_forward(assets, shape_params, expression_params, pose_params=None, eye_pose_params=None,
                  ignore_global_rot=False):
    """flame.py:180-217 with return_lm2d=return_lm3d=False: vertices [B,V,3].

    ``assets`` is the dict produced by oracle.synth.flame_assets (same tensors
    the reference registers as buffers at flame.py:74-90).
    """
    B = shape_params.shape[0]
    betas = torch.cat([shape_params, expression_params], 1)         # flame.py:192
    full_pose = flame_full_pose(pose_params, eye_pose_params, B, betas.dtype, ignore_global_rot)
    verts, _ = lbs(betas, full_pose, assets['v_template'], assets['shapedirs'], assets['posedirs'],
                   assets['J_regressor'], assets['parents'], assets['lbs_weights'], True)
    return verts


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary_coords):
    """lbs.py:102-138: barycentric landmarks.  lmk_faces_idx [B,L], bary [B,L,3]."""
    B = vertices.shape[0]
    tri = faces[lmk_faces_idx.reshape(-1)].view(B, -1, 3)            # [B,L,3] vertex ids
    idx = tri.unsqueeze(-1).expand(-1, -1, -1, 3)
    corners = torch.gather(vertices.unsqueeze(1).expand(-1, tri.shape[1], -1, -1), 2, idx)
    return torch.einsum('blfi,blf->bli', corners, lmk_bary_coords)
