"""EgoHMR.guide_coll / eval_coll restated with torch-CPU autograd (float64 by default).  TEST INFRASTRUCTURE (see
oracle/__init__).  The reference itself obtains this gradient with torch.autograd (models/egohmr/egohmr.py:517-570), so
the oracle does the same over its own restatement of de-normalise -> rot6d -> SMPL -> rotation_matrix_to_angle_axis;
the collision term is the caller's callable (COAP signature)."""
import numpy as np
import torch

UPPER_BODY = [0, 3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23]  # egohmr.py:567


class _Out:
    def __init__(self, vertices, joints, full_pose):
        self.vertices, self.joints, self.full_pose = vertices, joints, full_pose


def _normalize(v, eps=1e-12):
    return v / v.norm(dim=1, keepdim=True).clamp_min(eps)


def rot6d_to_rotmat(x):
    """utils/geometry.py:47-66, mode 'diffusion'."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = _normalize(a1)
    b2 = _normalize(a2 - (b1 * a2).sum(1, keepdim=True) * b1)
    b3 = torch.linalg.cross(b1, b2)
    return torch.stack((b1, b2, b3), dim=-1)


def _sdiv(num, den, eps=1e-6):
    den = den + (den.abs() < eps).to(den.dtype) * eps
    return num / den


def rotation_matrix_to_angle_axis(R, eps=1e-6):
    """utils/konia_transform.py:316-339,349-443,560-630 (differentiable; torch.where selects like the reference)."""
    m = R.reshape(-1, 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = (m[:, i] for i in range(9))
    trace = m00 + m11 + m22
    sq0 = torch.sqrt((trace + 1.0).clamp_min(eps)) * 2.0
    q0 = torch.stack([0.25 * sq0, _sdiv(m21 - m12, sq0), _sdiv(m02 - m20, sq0), _sdiv(m10 - m01, sq0)], -1)
    sq1 = torch.sqrt((1.0 + m00 - m11 - m22).clamp_min(eps)) * 2.0
    q1 = torch.stack([_sdiv(m21 - m12, sq1), 0.25 * sq1, _sdiv(m01 + m10, sq1), _sdiv(m02 + m20, sq1)], -1)
    sq2 = torch.sqrt((1.0 + m11 - m00 - m22).clamp_min(eps)) * 2.0
    q2 = torch.stack([_sdiv(m02 - m20, sq2), _sdiv(m01 + m10, sq2), 0.25 * sq2, _sdiv(m12 + m21, sq2)], -1)
    sq3 = torch.sqrt((1.0 + m22 - m00 - m11).clamp_min(eps)) * 2.0
    q3 = torch.stack([_sdiv(m10 - m01, sq3), _sdiv(m02 + m20, sq3), _sdiv(m12 + m21, sq3), 0.25 * sq3], -1)
    w2 = torch.where((m11 > m22)[:, None], q2, q3)
    w1 = torch.where(((m00 > m11) & (m00 > m22))[:, None], q1, w2)
    q = torch.where((trace > 0.0)[:, None], q0, w1)
    cos_t, a1, a2, a3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    s2 = a1 * a1 + a2 * a2 + a3 * a3
    sin_t = torch.sqrt(s2.clamp_min(eps))

    def safe_atan2(y, x):
        y = y + ((y.abs() < eps) & (x.abs() < eps)).to(y.dtype) * eps
        return torch.atan2(y, x)

    two_theta = 2.0 * torch.where(cos_t < 0.0, safe_atan2(-sin_t, -cos_t), safe_atan2(sin_t, cos_t))
    k = torch.where(s2 > 0.0, _sdiv(two_theta, sin_t), 2.0 * torch.ones_like(sin_t))
    return torch.stack([a1 * k, a2 * k, a3 * k], -1)


def smpl_forward(model, R, betas):
    """torch twin of oracle/smpl.py::smpl_forward (differentiable)."""
    dt = R.dtype
    B = R.shape[0]
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dt)
    v_t, S, P, Jr, W = t(model["v_template"]), t(model["shapedirs"]), t(model["posedirs"]), t(model["J_regressor"]), t(model["lbs_weights"])
    parents = [int(p) for p in model["parents"]]
    v_shaped = v_t[None] + torch.einsum("bl,mkl->bmk", betas, S)
    J = torch.einsum("bik,ji->bjk", v_shaped, Jr)
    pf = (R[:, 1:] - torch.eye(3, dtype=dt)).reshape(B, -1)
    v_posed = (pf @ P).reshape(B, -1, 3) + v_shaped
    rel = torch.cat([J[:, :1], J[:, 1:] - J[:, parents[1:]]], dim=1)
    Gr, Gt = [R[:, 0]], [rel[:, 0]]
    for i in range(1, 24):
        p = parents[i]
        Gr.append(Gr[p] @ R[:, i])
        Gt.append((Gr[p] @ rel[:, i, :, None])[..., 0] + Gt[p])
    Gr, Gt = torch.stack(Gr, 1), torch.stack(Gt, 1)
    At = Gt - (Gr @ J[..., None])[..., 0]
    Tr = torch.einsum("vj,bjrc->bvrc", W, Gr)
    Tt = torch.einsum("vj,bjr->bvr", W, At)
    verts = (Tr @ v_posed[..., None])[..., 0] + Tt
    joints = torch.cat([Gt, verts[:, [int(v) for v in model["extra_vertex_ids"]]]], dim=1)
    return verts, joints


def guide_coll(smpl_model, collision, x_t, betas, scene_pts, mean, std, dtype=torch.float64):
    """egohmr.py:517-570 -> grad [B,144] (numpy)."""
    x = torch.as_tensor(np.asarray(x_t), dtype=dtype).clone().requires_grad_()
    B = x.shape[0]
    pose = x * torch.as_tensor(np.asarray(std), dtype=dtype) + torch.as_tensor(np.asarray(mean), dtype=dtype)
    R = rot6d_to_rotmat(pose).reshape(B, 24, 3, 3)
    verts, joints = smpl_forward(smpl_model, R, torch.as_tensor(np.asarray(betas), dtype=dtype))
    aa = rotation_matrix_to_angle_axis(R.reshape(-1, 3, 3)).reshape(B, -1)
    pts_all = torch.as_tensor(np.asarray(scene_pts), dtype=dtype)
    losses = torch.zeros(B, dtype=dtype)
    for i in range(B):
        bb_min = verts[[i]].min(1).values.reshape(1, 3).detach()
        bb_max = verts[[i]].max(1).values.reshape(1, 3).detach()
        pts = pts_all[[i]]
        inds = (pts >= bb_min).all(-1) & (pts <= bb_max).all(-1)
        if inds.any():
            losses[i] = collision.collision_loss(pts[inds].unsqueeze(0), _Out(verts[[i]], joints[[i]], aa[[i]]))
    if int((losses == 0).sum()) >= B:
        return np.zeros((B, 144))
    # Reference quirk (egohmr.py:523-528,562): `x_t` is rebound to x_t*std+mean before autograd.grad(..., [x_t]), so
    # the gradient is taken w.r.t. the DE-NORMALISED pose (no std factor).
    g = torch.autograd.grad([-losses.mean()], [pose])[0].reshape(-1, 24, 6).clone()
    g[:, 3:] = g[:, 3:] * 2
    g[:, UPPER_BODY] = 0
    return g.reshape(-1, 144).numpy()
