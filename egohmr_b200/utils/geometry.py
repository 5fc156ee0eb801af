"""utils/geometry.py of the reference, for the functions the sampling path touches."""
import torch

_default_engine = {}


def _engine_for(device):
    from ..engine import Engine
    idx = device.index or 0
    if idx not in _default_engine:
        _default_engine[idx] = Engine(idx)
    return _default_engine[idx]


def rot6d_to_rotmat(x, rot6d_mode="prohmr", engine=None):
    """utils/geometry.py:47-66 on the GPU (K4).  x: [..., 6k] CUDA tensor -> [N, 3, 3].  Default mode 'prohmr' like the
    reference; every EgoHMR call site passes 'diffusion' explicitly (egohmr.py:260,529)."""
    if rot6d_mode == "prohmr":
        x = x.reshape(-1, 2, 3).permute(0, 2, 1)
    elif rot6d_mode != "diffusion":
        raise ValueError(f"rot6d_mode must be 'prohmr' or 'diffusion', got {rot6d_mode!r}")
    x6 = x.reshape(-1, 6).float().contiguous()
    if not x6.is_cuda:
        raise RuntimeError("rot6d_to_rotmat: CUDA tensor required (no CPU fallback)")
    return (engine or _engine_for(x6.device)).rot6d_to_rotmat(x6)


def batch_rodrigues(rot_vecs):
    """smplx/lbs.py::batch_rodrigues (what SMPL.forward(pose2rot=True) applies to axis-angle input — ground-truth bodies
    in compute_loss, egohmr.py:344-347, and in the driver, test_egohmr.py:307-312): angle = |v + 1e-8|,
    R = I + sin(angle) K + (1 - cos(angle)) K K."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos, sin = torch.cos(angle).unsqueeze(1), torch.sin(angle).unsqueeze(1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros_like(rx)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(-1, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def aa_to_rotmat(theta):
    """utils/geometry.py:5-43 (axis-angle -> quaternion -> matrix): the conversion compute_loss applies to the
    ground-truth pose for the parameter losses (egohmr.py:381)."""
    norm = torch.norm(theta + 1e-8, p=2, dim=1)
    angle = norm.unsqueeze(-1)
    normalized = theta / angle
    angle = angle * 0.5
    quat = torch.cat([torch.cos(angle), torch.sin(angle) * normalized], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz, 2 * wz + 2 * xy, w2 - x2 + y2 - z2,
                        2 * yz - 2 * wx, 2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def perspective_projection(points, translation, focal_length, camera_center=None, rotation=None):
    """utils/geometry.py:78-116 (identity rotation unless given).  Only the final step's 45 joints go through here."""
    if rotation is not None:
        points = torch.einsum("bij,bkj->bki", rotation, points)
    p = points + translation.unsqueeze(1)
    proj = p / p[:, :, -1:]
    if camera_center is None:
        camera_center = torch.zeros_like(focal_length)
    return proj[:, :, :2] * focal_length.unsqueeze(1) + camera_center.unsqueeze(1) * proj[:, :, 2:3]
