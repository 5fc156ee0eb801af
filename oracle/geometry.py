"""Rotation / projection helpers (numpy).  TEST INFRASTRUCTURE (see oracle/__init__)."""
import numpy as np


def _normalize(v, eps=1e-12):
    """F.normalize(v, dim=1): v / max(||v||_2, eps)."""
    n = np.sqrt((v * v).sum(axis=1, keepdims=True))
    return v / np.maximum(n, v.dtype.type(eps))


def rot6d_to_rotmat(x, rot6d_mode="diffusion"):
    """utils/geometry.py:47-66.  x: [..., 6k] -> [N, 3, 3] with columns (b1, b2, b3)."""
    if rot6d_mode == "prohmr":
        x = np.ascontiguousarray(x.reshape(-1, 2, 3).transpose(0, 2, 1))
    elif rot6d_mode == "diffusion":
        x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = _normalize(a1)
    b2 = _normalize(a2 - (b1 * a2).sum(axis=1, keepdims=True) * b1)
    b3 = np.cross(b1, b2)
    return np.stack((b1, b2, b3), axis=-1)


def rotmat_to_rot6d(R):
    """utils/geometry.py:69-75, mode 'diffusion': the first two columns, row-major."""
    return R[:, :, :-1].reshape(-1, 6)


def perspective_projection(points, translation, focal_length, camera_center):
    """utils/geometry.py:78-116 with identity rotation.  points [B,N,3], translation [B,3], focal/center [B,2]."""
    p = points + translation[:, None, :]
    proj = p / p[:, :, -1:]
    out = np.empty(points.shape[:2] + (2,), dtype=points.dtype)
    out[:, :, 0] = focal_length[:, None, 0] * proj[:, :, 0] + camera_center[:, None, 0] * proj[:, :, 2]
    out[:, :, 1] = focal_length[:, None, 1] * proj[:, :, 1] + camera_center[:, None, 1] * proj[:, :, 2]
    return out
