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


def rotation_matrix_to_angle_axis(R, eps=1e-6):
    """utils/konia_transform.py:316-339 = rotation_matrix_to_quaternion (:349-443, WXYZ) + quaternion_to_angle_axis
    (:560-630), including safe_zero_division (:343-346) and torch_safe_atan2 (:44-47).  R: [N,3,3] -> [N,3]."""
    dt = R.dtype
    m = R.reshape(-1, 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = (m[:, i] for i in range(9))
    trace = m00 + m11 + m22

    def sdiv(num, den):
        den = den.copy()
        den[np.abs(den) < eps] += dt.type(eps)
        return num / den

    def branch(sq_arg, order):
        sq = np.sqrt(np.maximum(sq_arg, dt.type(eps))) * dt.type(2.0)
        return sq, order(sq)

    sq0 = np.sqrt(np.maximum(trace + 1.0, eps)) * 2.0
    q0 = np.stack([0.25 * sq0, sdiv(m21 - m12, sq0), sdiv(m02 - m20, sq0), sdiv(m10 - m01, sq0)], -1)
    sq1 = np.sqrt(np.maximum(1.0 + m00 - m11 - m22, eps)) * 2.0
    q1 = np.stack([sdiv(m21 - m12, sq1), 0.25 * sq1, sdiv(m01 + m10, sq1), sdiv(m02 + m20, sq1)], -1)
    sq2 = np.sqrt(np.maximum(1.0 + m11 - m00 - m22, eps)) * 2.0
    q2 = np.stack([sdiv(m02 - m20, sq2), sdiv(m01 + m10, sq2), 0.25 * sq2, sdiv(m12 + m21, sq2)], -1)
    sq3 = np.sqrt(np.maximum(1.0 + m22 - m00 - m11, eps)) * 2.0
    q3 = np.stack([sdiv(m10 - m01, sq3), sdiv(m02 + m20, sq3), sdiv(m12 + m21, sq3), 0.25 * sq3], -1)
    w2 = np.where((m11 > m22)[:, None], q2, q3)
    w1 = np.where(((m00 > m11) & (m00 > m22))[:, None], q1, w2)
    q = np.where((trace > 0.0)[:, None], q0, w1).astype(dt)
    cos_t, a1, a2, a3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    s2 = a1 * a1 + a2 * a2 + a3 * a3
    sin_t = np.sqrt(np.maximum(s2, dt.type(eps)))

    def safe_atan2(y, x):
        y = y.copy()
        y[(np.abs(y) < eps) & (np.abs(x) < eps)] += dt.type(eps)
        return np.arctan2(y, x)

    two_theta = 2.0 * np.where(cos_t < 0.0, safe_atan2(-sin_t, -cos_t), safe_atan2(sin_t, cos_t))
    k = np.where(s2 > 0.0, sdiv(two_theta, sin_t), 2.0 * np.ones_like(sin_t))
    return np.stack([a1 * k, a2 * k, a3 * k], -1).astype(dt)
