"""Host-side logic that needs no GPU: schedule tables and per-step coefficients of the product sampler, the C-ABI
library's exported symbols, and the image-sharding plan."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from egohmr_b200 import _lib
from egohmr_b200.diffusion.model_util import create_gaussian_diffusion
from egohmr_b200.diffusion.respace import space_timesteps
from oracle import schedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("T,resp", [(50, ""), (50, "ddim5"), (100, ""), (1000, "ddim5"), (1000, "")])
def test_product_tables_match_reference_golden(golden_dir, T, resp):
    tab = np.load(os.path.join(golden_dir, "schedule_tables.npz"))
    d = create_gaussian_diffusion(T, resp)
    tag = f"T{T}_{resp or 'ddpm'}"
    assert d.timestep_map == list(tab[tag + "_timestep_map"])
    for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"):
        assert np.array_equal(getattr(d, k), tab[f"{tag}_{k}"]), k


def test_space_timesteps_matches_oracle_and_errors():
    for T in (50, 100, 1000):
        for s in ("ddim5", "ddim10", [T], "10,15,16", "3", "1", "7,9"):
            assert space_timesteps(T, s) == schedule.space_timesteps(T, s)
    with pytest.raises(ValueError):
        space_timesteps(50, "ddim11")     # no integer stride gives 11 steps of 50
    with pytest.raises(ValueError):
        space_timesteps(10, "20")         # cannot take 20 steps out of 10


def test_step_coefficients_match_oracle():
    d = create_gaussian_diffusion(50, "ddim5")
    s = schedule.Schedule(50, "ddim5")
    c = d.step_coefficients(d.DDIM)
    for t in range(5):
        ref = np.array(schedule.ddim_coefficients(s, t), np.float32)
        assert np.array_equal(c[t, :4], ref), (t, c[t, :4], ref)
    assert c[0, 2] == 1.0 and c[0, 3] == 0.0   # last DDIM step returns pred_xstart exactly
    d = create_gaussian_diffusion(100, "")
    s = schedule.Schedule(100, "")
    c = d.step_coefficients(d.DDPM, guided=True, cond_grad_weight=2.0)
    for t in range(100):
        ref = np.array(schedule.ddpm_coefficients(s, t, True, 2.0), np.float32)
        assert np.allclose(c[t, :4], ref, rtol=2e-7, atol=0), (t, c[t, :4], ref)  # expf may differ by 1 ulp
        assert c[t, 0] == ref[0] and c[t, 1] == ref[1]
    assert c[0, 2] == 0.0                      # no noise at t == 0
    assert (c[11:, 3] == 0).all() and (c[:11, 3] > 0).all()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "egohmr_b200.h")).read()
    declared = set(re.findall(r"\b(ehb_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("ehb_ctx")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    _lib.load()


def test_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/egohmr_b200.h must compile as C99 on its own (plain pointers and sizes, no
    C++ or torch types), and a C translation unit that takes the address of every entry point must link against the library."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    hdr = os.path.join(ROOT, "include", "egohmr_b200.h")
    subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", hdr], check=True)
    names = sorted(_lib.SIGNATURES)
    src = tmp_path / "abi.c"
    src.write_text('#include "egohmr_b200.h"\n#include <stdio.h>\nint main(void) {\n  void* f[] = {' +
                   ", ".join(f"(void*)&{n}" for n in names) + "};\n  printf(\"%d\\n\", (int)(sizeof f / sizeof f[0]));\n  return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L" + libdir,
                    "-legohmr_b200", "-Wl,-rpath," + libdir, "-Wl,--unresolved-symbols=ignore-in-shared-libs"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    # the executable loads the library (and libcudart behind it); where the CUDA runtime cannot be loaded only the link is checked
    if out.returncode == 0:
        assert int(out.stdout) == len(names)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from egohmr_b200.engine import Engine
    with pytest.raises(_lib.EhbError):
        Engine(0)


def test_preprocess_stats_loader(tmp_path):
    """test_egohmr.py:108-111: <logdir>/preprocess_stats/preprocess_stats.npz next to the checkpoint."""
    import numpy as np
    import pytest
    from egohmr_b200 import checkpoint
    d = tmp_path / "run" / "preprocess_stats"
    d.mkdir(parents=True)
    mean, std = np.arange(144, dtype=np.float64), np.full(144, 0.5)
    np.savez(d / "preprocess_stats.npz", Xmean=mean, Xstd=std)
    m, s = checkpoint.load_preprocess_stats(str(tmp_path / "run" / "best_model_mpjpe_vis.pt"))
    assert m.dtype.is_floating_point and m.shape == (144,) and float(m[5]) == 5.0 and float(s[0]) == 0.5
    np.savez(d / "preprocess_stats.npz", Xmean=mean[:10], Xstd=std[:10])
    with pytest.raises(ValueError):
        checkpoint.load_preprocess_stats(str(tmp_path / "run" / "x.pt"))
    np.savez(d / "preprocess_stats.npz", Xmean=mean)
    with pytest.raises(KeyError):
        checkpoint.load_preprocess_stats(str(tmp_path / "run" / "x.pt"))


def test_stage1_handoff_and_scene_cloud_readers(tmp_path):
    """`results.pkl['pred_cam_full_list']` (egobody_dataset.py:94-98) and the per-frame scene cloud `.npy`
    (:213-225, :270-273) in the formats the reference's dataloader reads."""
    import pickle
    from egohmr_b200 import data
    rng = np.random.default_rng(0)
    cams = rng.normal(0, 1, (10, 3))
    with open(tmp_path / "results.pkl", "wb") as fp:
        pickle.dump({"pred_cam_full_list": cams, "other": 1}, fp)
    t = data.load_stage1_translations(tmp_path / "results.pkl", spacing=3)
    assert t.dtype == np.float32 and t.shape == (4, 3) and np.array_equal(t, cams.astype(np.float64)[::3].astype(np.float32))
    with open(tmp_path / "bad.pkl", "wb") as fp:
        pickle.dump({"pred_cam": cams}, fp)
    with pytest.raises(KeyError):
        data.load_stage1_translations(tmp_path / "bad.pkl")
    pts = rng.normal(0, 1, (20000, 3))
    np.save(tmp_path / "frame_01234.npy", pts)
    T = np.eye(4)
    T[:3, :3] = np.linalg.qr(rng.normal(0, 1, (3, 3)))[0]
    T[:3, 3] = [0.1, -0.2, 3.0]
    got = data.load_scene_cloud(tmp_path / "frame_01234.npy", T, downsample_rate=2, n_points=10000)
    want = (pts @ T[:3, :3].T + T[:3, 3]).astype(np.float32)[::2]           # utils/geometry.py:137-141
    assert got.dtype == np.float32 and np.array_equal(got, want)
    with pytest.raises(ValueError):
        data.load_scene_cloud(tmp_path / "frame_01234.npy", T, n_points=5000)
    b = data.make_batch(np.zeros((2, 3, 224, 224), np.float32), np.ones((2, 25, 3)), [1500.0, 3000.0], [960, 960], [540, 540],
                        np.zeros((2, 2)), [300, 400], t[:2], [got, got], device="cpu")
    assert b["fx"].tolist() == [1.0, 2.0] and b["scene_pcd_verts_full"].shape == (2, 10000, 3)
    assert b["smpl_params"]["transl"] is b["stage1_transl_full"] and b["img"].dtype == torch.float32


def test_rescale_timesteps_follows_the_wrapped_model_order():
    """respace.py:111-129: SpacedDiffusion never scales the sampler index itself; the wrapped model maps the respaced index to
    the original timestep FIRST and then scales by 1000 / original_num_steps (the factory passes rescale_timesteps=False, so
    this only matters to callers that build SpacedDiffusion themselves)."""
    from egohmr_b200.diffusion.gaussian_diffusion import get_named_beta_schedule
    from egohmr_b200.diffusion.respace import SpacedDiffusion, space_timesteps
    betas = get_named_beta_schedule("cosine", 50)
    for rescale in (False, True):
        d = SpacedDiffusion(use_timesteps=space_timesteps(50, "ddim5"), betas=betas, rescale_timesteps=rescale)
        t = torch.tensor([0, 1, 4, 3])
        assert torch.equal(d._scale_timesteps(t), t)
        got = d._map_timesteps(t)
        want = torch.tensor(d.timestep_map)[t]
        if rescale:
            assert got.dtype == torch.float32 and torch.allclose(got, want.float() * (1000.0 / 50))
        else:
            assert torch.equal(got, want)
