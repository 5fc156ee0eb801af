"""Procrustes-aligned reconstruction error (numpy).  TEST INFRASTRUCTURE (see oracle/__init__).
Restates utils/pose_utils.py:11-125 of the reference; pinned by tests/golden/procrustes.npz (made by the reference)."""
import numpy as np


def compute_similarity_transform(S1, S2, mask=None, dtype=np.float64):
    """pose_utils.py:11-56 (and :72-105 when `mask` [N,3] is given).  S1, S2 [N,3] -> S1_hat [N,3]."""
    A = S1.astype(dtype)
    a, b = (A if mask is None else A * mask).T, (S2.astype(dtype) if mask is None else S2.astype(dtype) * mask).T
    mu1, mu2 = a.mean(axis=1, keepdims=True), b.mean(axis=1, keepdims=True)
    X1, X2 = a - mu1, b - mu2
    var1 = np.sum(X1 ** 2)
    K = X1.dot(X2.T)
    U, s, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(3)
    Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
    R = V.dot(Z.dot(U.T))
    scale = np.trace(R.dot(K)) / var1
    t = mu2 - scale * R.dot(mu1)
    return (scale * R.dot(A.T) + t).T


def reconstruction_error(S1, S2, mask=None, avg_joint=True, dtype=np.float64):
    """pose_utils.py:108-125."""
    hat = np.stack([compute_similarity_transform(S1[i], S2[i], None if mask is None else mask[i], dtype)
                    for i in range(S1.shape[0])])
    re = np.sqrt(((hat - S2.astype(dtype)) ** 2).sum(axis=-1))
    return (re.mean(axis=-1) if avg_joint else re), hat


def nn_dist_sq(x, y):
    """Squared distance of every point of x [N,P1,3] to its nearest point of y [N,P2,3]: what
    utils/pytorch3d_chamfer_distance.py:160-164 gets from pytorch3d's knn_points(K=1).dists[..., 0].  PARITY UNPINNED:
    pytorch3d is a third-party dependency that is not vendored in the reference and not installed here; this is the
    definition (brute force), evaluated in float64."""
    x, y = x.astype(np.float64), y.astype(np.float64)
    out = np.empty(x.shape[:2])
    for n in range(x.shape[0]):
        d = ((x[n][:, None, :] - y[n][None, :, :]) ** 2).sum(-1)
        out[n] = d.min(axis=1)
    return out
