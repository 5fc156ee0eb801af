"""Known-answer tests for the (parity-unpinned) SMPL restatement, oracle/smpl.py."""
import numpy as np

from egohmr_b200 import synth
from oracle import geometry, smpl


def _rot(axis, angle):
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def test_identity_pose_zero_betas_is_template():
    m = synth.make_smpl_model(0, n_verts=500)
    R = np.tile(np.eye(3), (2, 24, 1, 1))
    out = smpl.smpl_forward(m, R, np.zeros((2, 10)))
    # skinning weights are normalised and then stored in fp32, so they sum to 1 only to ~1e-7
    assert np.abs(out["vertices"] - m["v_template"][None]).max() < 1e-6
    J = m["J_regressor"].astype(np.float64) @ m["v_template"].astype(np.float64)
    assert np.abs(out["joints"][:, :24] - J[None]).max() < 1e-12


def test_root_rotation_is_rigid_about_root_joint():
    m = synth.make_smpl_model(1, n_verts=500)
    betas = np.random.default_rng(0).normal(0, 1, (1, 10))
    R = np.tile(np.eye(3), (1, 24, 1, 1))
    base = smpl.smpl_forward(m, R, betas)
    Rg = _rot([0.3, 1.0, -0.2], 0.7)
    R2 = R.copy()
    R2[0, 0] = Rg
    out = smpl.smpl_forward(m, R2, betas)
    J0 = base["joints"][0, 0]
    expect = (base["vertices"][0] - J0) @ Rg.T + J0   # pose blend shapes ignore the root (pose_feature uses R[1:])
    assert np.abs(out["vertices"][0] - expect).max() < 1e-6


def test_translation_and_extra_joints():
    m = synth.make_smpl_model(2, n_verts=500)
    rng = np.random.default_rng(1)
    R = geometry.rot6d_to_rotmat(rng.normal(0, 1, (3, 144))).reshape(3, 24, 3, 3)
    betas = rng.normal(0, 1, (3, 10))
    t = rng.normal(0, 1, (3, 3))
    a = smpl.smpl_forward(m, R, betas)
    b = smpl.smpl_forward(m, R, betas, t)
    assert np.abs(b["vertices"] - a["vertices"] - t[:, None]).max() < 1e-12
    assert b["joints"].shape == (3, 45, 3)
    assert np.abs(b["joints"][:, 24:] - b["vertices"][:, m["extra_vertex_ids"]]).max() == 0
