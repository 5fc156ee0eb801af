"""ctypes binding of libegohmr_b200.so (include/egohmr_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (``egohmr_b200/lib/libegohmr_b200.so``).  There is no
fallback: if the shared object is missing, or a context cannot be created because there is no sm_100 GPU, the caller
gets an exception — never a silent CPU / eager-PyTorch path.
"""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libegohmr_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class EhbError(RuntimeError):
    pass


class GConv(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("out_dim", C.c_int32), ("W", c_float_p), ("M", c_float_p), ("adj2", c_float_p),
                ("bias", c_float_p), ("bn_weight", c_float_p), ("bn_bias", c_float_p), ("bn_mean", c_float_p),
                ("bn_var", c_float_p), ("bn_eps", C.c_float)]


class GcnWeights(C.Structure):
    _fields_ = [("hid", C.c_int32), ("n_blocks", C.c_int32), ("img_dim", C.c_int32), ("cond_dim", C.c_int32),
                ("xfeat_dim", C.c_int32), ("temb_dim", C.c_int32), ("diffuse_fuse", C.c_int32), ("adj", c_float_p),
                ("inproc_w", c_float_p), ("inproc_b", c_float_p), ("layers", C.POINTER(GConv)), ("n_layers", C.c_int32),
                ("mask_all_cond", C.c_int32)]


class NonLocalWeights(C.Structure):
    _fields_ = [("inter", C.c_int32)] + [(n, c_float_p) for n in
                                         ("theta_w", "theta_b", "phi_w", "phi_b", "g_w", "g_b", "W_w", "W_b", "bn_weight",
                                          "bn_bias", "bn_mean", "bn_var")] + [("bn_eps", C.c_float)]


class ConvBN(C.Structure):
    _fields_ = [("cout", C.c_int32), ("cin", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32),
                ("pad", C.c_int32), ("weight", c_float_p), ("bn_weight", c_float_p), ("bn_bias", c_float_p),
                ("bn_mean", c_float_p), ("bn_var", c_float_p), ("bn_eps", C.c_float)]


class ResnetWeights(C.Structure):
    _fields_ = [("convs", C.POINTER(ConvBN)), ("n_convs", C.c_int32), ("blocks", C.c_int32 * 4)]


class PointnetWeights(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("out_dim", C.c_int32), ("fc_pos_w", c_float_p), ("fc_pos_b", c_float_p),
                ("fc0_w", c_float_p * 4), ("fc0_b", c_float_p * 4), ("fc1_w", c_float_p * 4), ("fc1_b", c_float_p * 4),
                ("shortcut_w", c_float_p * 4), ("fc_c_w", c_float_p), ("fc_c_b", c_float_p)]


class SmplModel(C.Structure):
    _fields_ = [("n_verts", C.c_int32), ("n_betas", C.c_int32), ("n_extra", C.c_int32), ("v_template", c_float_p),
                ("shapedirs", c_float_p), ("posedirs", c_float_p), ("J_regressor", c_float_p),
                ("lbs_weights", c_float_p), ("parents", c_int32_p), ("extra_vertex_ids", c_int32_p)]


# every symbol include/egohmr_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SIGNATURES = {
    "ehb_last_error": (C.c_char_p, []),
    "ehb_launch_count": (C.c_int64, [_vp]),
    "ehb_alloc_epoch": (C.c_uint64, []),
    "ehb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ehb_ctx_destroy": (None, [_vp]),
    "ehb_gcn_load": (C.c_int, [_vp, C.POINTER(GcnWeights)]),
    "ehb_gcn_load_nonlocal": (C.c_int, [_vp, C.POINTER(NonLocalWeights)]),
    "ehb_smpl_load": (C.c_int, [_vp, C.POINTER(SmplModel)]),
    "ehb_set_norm": (C.c_int, [_vp, c_float_p, c_float_p]),
    "ehb_set_schedule": (C.c_int, [_vp, C.c_int, C.c_int, c_float_p]),
    "ehb_set_cond": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "ehb_set_temb": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "ehb_set_bodies": (C.c_int, [_vp, C.c_int, c_int32_p]),
    "ehb_denoise_step": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_denoise_step_ex": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_denoise_step_debug": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_sampler_update": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_sampler_update_ex": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_rot6d_to_rotmat": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "ehb_decode": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_smpl_forward": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_pointnet_load": (C.c_int, [_vp, C.POINTER(PointnetWeights)]),
    "ehb_pointnet_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "ehb_maxpool3x3s2_nhwc": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "ehb_cond_inputs": (C.c_int, [_vp] + [_vp] * 9 + [C.c_int] * 7 + [c_int32_p, C.c_float, _vp, _vp, _vp, _vp]),
    "ehb_project_joints": (C.c_int, [_vp] + [_vp] * 6 + [C.c_int, C.c_int, C.c_float, C.c_float] + [_vp] * 5),
    "ehb_scene_crop": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "ehb_procrustes_align": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "ehb_eval_metrics": (C.c_int, [_vp] + [_vp] * 8 + [C.c_int] * 4 + [_vp] * 5),
    "ehb_resnet_load": (C.c_int, [_vp, C.POINTER(ResnetWeights)]),
    "ehb_resnet_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "ehb_linear_f32": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "ehb_nn_dist_sq": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "ehb_rotmat_to_angle_axis": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "ehb_smpl_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ehb_debug_set_gemm_mode": (C.c_int, [_vp, C.c_int]),
    "ehb_debug_set_resnet_mode": (C.c_int, [_vp, C.c_int]),
    "ehb_debug_set_pdl": (C.c_int, [_vp, C.c_int]),
    "ehb_debug_set_input_mode": (C.c_int, [_vp, C.c_int]),
    "ehb_debug_set_k1_fused": (C.c_int, [_vp, C.c_int]),
    "ehb_debug_set_conv_kc": (C.c_int, [_vp, C.c_int]),
    "ehb_debug_gemm_hl": (C.c_int, [_vp, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                    c_float_p]),
    "ehb_check_overflow": (C.c_int, [_vp, _vp]),
    "ehb_overflow_flag_async": (C.c_int, [_vp, _vp, _vp]),
    "ehb_time_hidden_layer": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_float), _vp]),
    "ehb_time_stage": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_float), _vp]),
    "ehb_debug_get_buffer": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """dlopen the in-tree library and type every exported entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EhbError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` first "
                       "(there is no CPU / PyTorch fallback for the sampling hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EhbError(load().ehb_last_error().decode("utf-8", "replace"))


def f32(a):
    """Contiguous float32 numpy view/copy (kept alive by the caller for the duration of the call)."""
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def fptr(a):
    return a.ctypes.data_as(c_float_p)
