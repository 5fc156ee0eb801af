"""Thin Python handle on one ``ehb_ctx`` (one per process / GPU).  PyTorch is only the owner of device memory and of
the CUDA stream here; all arithmetic of the hot path happens inside libegohmr_b200.so."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, f32, fptr


def _dev_ptr(t, dtype=torch.float32, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError("tensor required")
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"expected a contiguous CUDA {dtype} tensor, got "
                         f"{type(t).__name__} {getattr(t, 'dtype', None)} cuda={getattr(t, 'is_cuda', None)}")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    def __init__(self, device=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.EhbError("egohmr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = C.c_void_p()
        check(self.lib.ehb_ctx_create(self.device.index, C.byref(h)))
        self._h = h
        self.n_bodies = 0
        self.n_img = 0
        self.n_verts = 0
        self.n_extra = 0
        self.n_betas = 0
        self.hid = 0
        self.smpl_loaded = False
        self.gcn_loaded = False

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ehb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ loading
    def load_gcn(self, sd, adj, hid, n_blocks, diffuse_fuse=True, img_dim=2048, cond_dim=2694, xfeat_dim=512,
                 temb_dim=512, prefix="diffusion_model", bn_eps=1e-5, mask_all_cond=False):
        """`sd`: mapping reference-state_dict-name -> array (numpy or torch), see include/egohmr_b200.h::ehb_gconv."""
        keep = []

        def arr(name):
            v = sd[name]
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            a = f32(v)
            keep.append(a)
            return a

        def gconv(gname, bnname, in_dim, out_dim):
            g = _lib.GConv()
            g.in_dim, g.out_dim = in_dim, out_dim
            W = arr(gname + ".W")
            assert W.shape == (2, in_dim, out_dim), (gname, W.shape)
            g.W, g.M, g.adj2, g.bias = fptr(W), fptr(arr(gname + ".M")), fptr(arr(gname + ".adj2")), fptr(arr(gname + ".bias"))
            if bnname is not None:
                g.bn_weight, g.bn_bias = fptr(arr(bnname + ".weight")), fptr(arr(bnname + ".bias"))
                g.bn_mean, g.bn_var = fptr(arr(bnname + ".running_mean")), fptr(arr(bnname + ".running_var"))
            g.bn_eps = bn_eps
            return g

        in_dim = cond_dim + xfeat_dim + temb_dim
        layers = [gconv(f"{prefix}.gconv_input.0.gconv", f"{prefix}.gconv_input.0.bn", in_dim, hid)]
        for b in range(n_blocks):
            for k in (1, 2):
                layers.append(gconv(f"{prefix}.gconv_layers.{b}.gconv{k}.gconv", f"{prefix}.gconv_layers.{b}.gconv{k}.bn",
                                    hid, hid))
        layers.append(gconv(f"{prefix}.gconv_output", None, hid, 6))
        arr_t = (_lib.GConv * len(layers))(*layers)
        w = _lib.GcnWeights()
        w.hid, w.n_blocks, w.img_dim, w.cond_dim, w.xfeat_dim, w.temb_dim = hid, n_blocks, img_dim, cond_dim, xfeat_dim, temb_dim
        w.diffuse_fuse = 1 if diffuse_fuse else 0
        w.mask_all_cond = 1 if mask_all_cond else 0
        adj_a = f32(adj.detach().cpu().numpy() if isinstance(adj, torch.Tensor) else adj)
        w.adj = fptr(adj_a)
        w.inproc_w, w.inproc_b = fptr(arr("input_process.poseEmbedding.weight")), fptr(arr("input_process.poseEmbedding.bias"))
        w.layers, w.n_layers = arr_t, len(layers)
        check(self.lib.ehb_gcn_load(self._h, C.byref(w)))
        self.hid, self.gcn_loaded, self.n_bodies = hid, True, 0

    def load_nonlocal(self, sd, prefix="diffusion_model.non_local", bn_eps=1e-5):
        """NONLocalBlock2D parameters by their reference names (nets/non_local_embedded_gaussian.py:36-54); `sd=None`
        switches the block off."""
        if sd is None:
            check(self.lib.ehb_gcn_load_nonlocal(self._h, None))
            return
        keep = []

        def arr(name):
            v = sd[f"{prefix}.{name}"]
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            a = f32(np.asarray(v).reshape(v.shape[0], -1) if np.asarray(v).ndim > 1 else v)
            keep.append(a)
            return fptr(a)

        w = _lib.NonLocalWeights()
        w.inter = int(sd[f"{prefix}.theta.weight"].shape[0])
        w.theta_w, w.theta_b = arr("theta.weight"), arr("theta.bias")
        w.phi_w, w.phi_b = arr("phi.weight"), arr("phi.bias")
        w.g_w, w.g_b = arr("g.weight"), arr("g.bias")
        w.W_w, w.W_b = arr("W.0.weight"), arr("W.0.bias")
        w.bn_weight, w.bn_bias = arr("W.1.weight"), arr("W.1.bias")
        w.bn_mean, w.bn_var = arr("W.1.running_mean"), arr("W.1.running_var")
        w.bn_eps = bn_eps
        check(self.lib.ehb_gcn_load_nonlocal(self._h, C.byref(w)))

    def load_smpl(self, model):
        a = {k: f32(model[k]) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
        parents = np.ascontiguousarray(model["parents"], dtype=np.int32)
        extra = np.ascontiguousarray(model["extra_vertex_ids"], dtype=np.int32)
        m = _lib.SmplModel()
        m.n_verts, m.n_betas, m.n_extra = a["v_template"].shape[0], a["shapedirs"].shape[-1], extra.shape[0]
        assert a["shapedirs"].shape == (m.n_verts, 3, m.n_betas) and a["posedirs"].shape == (207, m.n_verts * 3)
        m.v_template, m.shapedirs, m.posedirs = fptr(a["v_template"]), fptr(a["shapedirs"]), fptr(a["posedirs"])
        m.J_regressor, m.lbs_weights = fptr(a["J_regressor"]), fptr(a["lbs_weights"])
        m.parents = parents.ctypes.data_as(_lib.c_int32_p)
        m.extra_vertex_ids = extra.ctypes.data_as(_lib.c_int32_p)
        check(self.lib.ehb_smpl_load(self._h, C.byref(m)))
        self.n_verts, self.n_extra, self.n_betas, self.smpl_loaded = m.n_verts, m.n_extra, m.n_betas, True

    def load_resnet(self, sd, prefix="backbone", blocks=(3, 4, 6, 3), bn_eps=1e-5):
        """ResNet-50 parameters by their reference names (models/resnet.py:100-150) -> tcgen05 convolution GEMMs."""
        keep, convs = [], []

        def arr(name):
            v = sd[f"{prefix}.{name}"]
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            a = f32(v)
            keep.append(a)
            return a

        def conv(cname, bname, stride, pad):
            w = arr(cname + ".weight")
            c = _lib.ConvBN()
            c.cout, c.cin, c.kh, c.kw = (int(x) for x in w.shape)
            c.stride, c.pad = stride, pad
            c.weight = fptr(w)
            c.bn_weight, c.bn_bias = fptr(arr(bname + ".weight")), fptr(arr(bname + ".bias"))
            c.bn_mean, c.bn_var = fptr(arr(bname + ".running_mean")), fptr(arr(bname + ".running_var"))
            c.bn_eps = bn_eps
            convs.append(c)

        conv("conv1", "bn1", 2, 3)
        for li, nb in enumerate(blocks, start=1):
            for bi in range(nb):
                p = f"layer{li}.{bi}"
                s = (1 if li == 1 else 2) if bi == 0 else 1
                conv(p + ".conv1", p + ".bn1", 1, 0)
                conv(p + ".conv2", p + ".bn2", s, 1)
                conv(p + ".conv3", p + ".bn3", 1, 0)
                if bi == 0:
                    conv(p + ".downsample.0", p + ".downsample.1", s, 0)
        w = _lib.ResnetWeights()
        arr_t = (_lib.ConvBN * len(convs))(*convs)
        w.convs, w.n_convs = arr_t, len(convs)
        for i, nb in enumerate(blocks):
            w.blocks[i] = nb
        check(self.lib.ehb_resnet_load(self._h, C.byref(w)))
        self.resnet_out = int(convs[-1].cout)

    def resnet_forward(self, img):
        """backbone(img): [n,3,H,W] fp32 NCHW -> [n, 2048] (K9, tcgen05 convolution GEMMs, fp32-class)."""
        n, _, h, w = img.shape
        out = torch.empty(n, self.resnet_out, device=img.device, dtype=torch.float32)
        check(self.lib.ehb_resnet_forward(self._h, _dev_ptr(img), n, h, w, _dev_ptr(out), _stream()))
        return out

    def load_pointnet(self, sd, prefix="scene_enc", hidden=256):
        """ResnetPointnet parameters by their reference names (models/respointnet.py:13-27)."""
        keep = []

        def arr(name):
            v = sd[f"{prefix}.{name}"]
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            a = f32(v)
            keep.append(a)
            return fptr(a)

        w = _lib.PointnetWeights()
        w.hidden = hidden
        w.out_dim = int(sd[f"{prefix}.fc_c.weight"].shape[0])
        w.fc_pos_w, w.fc_pos_b = arr("fc_pos_0.weight"), arr("fc_pos_0.bias")
        for i in range(4):
            w.fc0_w[i], w.fc0_b[i] = arr(f"block_{i}.fc_0.weight"), arr(f"block_{i}.fc_0.bias")
            w.fc1_w[i], w.fc1_b[i] = arr(f"block_{i}.fc_1.weight"), arr(f"block_{i}.fc_1.bias")
            w.shortcut_w[i] = arr(f"block_{i}.shortcut.weight")
        w.fc_c_w, w.fc_c_b = arr("fc_c.weight"), arr("fc_c.bias")
        check(self.lib.ehb_pointnet_load(self._h, C.byref(w)))
        self.pointnet_out = w.out_dim

    def pointnet_forward(self, pts):
        """scene_enc(pts): [n_clouds, n_pts, 3] -> [n_clouds, out_dim] (K7, tcgen05)."""
        n_clouds, n_pts = pts.shape[0], pts.shape[1]
        out = torch.empty(n_clouds, self.pointnet_out, device=pts.device, dtype=torch.float32)
        check(self.lib.ehb_pointnet_forward(self._h, _dev_ptr(pts), n_clouds, n_pts, _dev_ptr(out), _stream()))
        return out

    def maxpool3x3s2(self, x):
        """nn.MaxPool2d(3, 2, 1) on a channels_last [N,C,H,W] fp32 CUDA tensor (NHWC in memory) -> channels_last."""
        N, Cc, H, W = x.shape
        if not x.is_contiguous(memory_format=torch.channels_last) or x.dtype != torch.float32 or not x.is_cuda:
            raise ValueError("maxpool3x3s2 expects a channels_last float32 CUDA tensor")
        out = torch.empty((N, Cc, (H + 1) // 2, (W + 1) // 2), device=x.device, dtype=torch.float32,
                          memory_format=torch.channels_last)
        check(self.lib.ehb_maxpool3x3s2_nhwc(self._h, C.c_void_p(x.data_ptr()), N, H, W, Cc, C.c_void_p(out.data_ptr()),
                                             _stream()))
        return out

    def linear(self, x, w_t, bias=None, relu=False):
        """nn.Linear(+ReLU) in the library's fp32 FFMA GEMM: x [M,K], w_t [K,N] (transposed weight) -> [M,N]."""
        M, K = x.shape
        N = w_t.shape[1]
        y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        check(self.lib.ehb_linear_f32(self._h, _dev_ptr(x), _dev_ptr(w_t), _dev_ptr(bias, allow_none=True), M, N, K,
                                      1 if relu else 0, _dev_ptr(y), _stream()))
        return y

    def set_norm(self, mean, std):
        m, s = f32(mean).reshape(144), f32(std).reshape(144)
        check(self.lib.ehb_set_norm(self._h, fptr(m), fptr(s)))

    def set_schedule(self, kind, coef):
        coef = f32(coef).reshape(-1, 8)
        check(self.lib.ehb_set_schedule(self._h, int(kind), coef.shape[0], fptr(coef)))

    # ------------------------------------------------------------------ per batch
    def set_cond(self, img_feat, rest_feat, vis):
        n = img_feat.shape[0]
        check(self.lib.ehb_set_cond(self._h, n, _dev_ptr(img_feat), _dev_ptr(rest_feat), _dev_ptr(vis, torch.uint8),
                                    _stream()))
        self.n_img = n

    def set_temb(self, temb):
        check(self.lib.ehb_set_temb(self._h, temb.shape[0], _dev_ptr(temb), _stream()))

    def set_bodies(self, img_of_body):
        iob = np.ascontiguousarray(img_of_body, dtype=np.int32)
        check(self.lib.ehb_set_bodies(self._h, iob.shape[0], iob.ctypes.data_as(_lib.c_int32_p)))
        self.n_bodies = iob.shape[0]

    # ------------------------------------------------------------------ hot calls
    def denoise_step(self, step, x_t, noise, grad, x_prev, x0, out_cond=None, out_uncond=None, x0_model=None):
        if x0_model is not None:
            check(self.lib.ehb_denoise_step_ex(self._h, int(step), _dev_ptr(x_t), _dev_ptr(noise, allow_none=True),
                                               _dev_ptr(grad, allow_none=True), _dev_ptr(x_prev), _dev_ptr(x0),
                                               _dev_ptr(x0_model), _stream()))
        elif out_cond is None and out_uncond is None:
            check(self.lib.ehb_denoise_step(self._h, int(step), _dev_ptr(x_t), _dev_ptr(noise, allow_none=True),
                                            _dev_ptr(grad, allow_none=True), _dev_ptr(x_prev), _dev_ptr(x0), _stream()))
        else:
            check(self.lib.ehb_denoise_step_debug(self._h, int(step), _dev_ptr(x_t), _dev_ptr(noise, allow_none=True),
                                                  _dev_ptr(grad, allow_none=True), _dev_ptr(x_prev), _dev_ptr(x0),
                                                  _dev_ptr(out_cond, allow_none=True),
                                                  _dev_ptr(out_uncond, allow_none=True), _stream()))

    def sampler_update(self, step, x_t, x0, noise, grad, x_prev, x0_out=None):
        check(self.lib.ehb_sampler_update_ex(self._h, int(step), x_t.shape[0], _dev_ptr(x_t), _dev_ptr(x0),
                                             _dev_ptr(noise, allow_none=True), _dev_ptr(grad, allow_none=True),
                                             _dev_ptr(x_prev), _dev_ptr(x0_out, allow_none=True), _stream()))

    def decode(self, x0, betas, want_smpl=True):
        """x0 [B,144] (normalised) -> pose6d [B,144], R [B,24,3,3], verts [B,V,3] | None, joints [B,24+E,3] | None."""
        B = x0.shape[0]
        assert B == self.n_bodies
        pose6d = torch.empty(B, 144, device=x0.device, dtype=torch.float32)
        R = torch.empty(B, 24, 3, 3, device=x0.device, dtype=torch.float32)
        verts = joints = None
        if want_smpl:
            verts = torch.empty(B, self.n_verts, 3, device=x0.device, dtype=torch.float32)
            joints = torch.empty(B, 24 + self.n_extra, 3, device=x0.device, dtype=torch.float32)
        check(self.lib.ehb_decode(self._h, _dev_ptr(x0), _dev_ptr(betas, allow_none=not want_smpl), _dev_ptr(pose6d),
                                  _dev_ptr(R), _dev_ptr(verts, allow_none=True), _dev_ptr(joints, allow_none=True),
                                  _stream()))
        return pose6d, R, verts, joints

    def smpl_forward(self, R, betas, transl=None):
        n = R.shape[0]
        verts = torch.empty(n, self.n_verts, 3, device=R.device, dtype=torch.float32)
        joints = torch.empty(n, 24 + self.n_extra, 3, device=R.device, dtype=torch.float32)
        check(self.lib.ehb_smpl_forward(self._h, n, _dev_ptr(R), _dev_ptr(betas), _dev_ptr(transl, allow_none=True),
                                        _dev_ptr(verts), _dev_ptr(joints), _stream()))
        return verts, joints

    def rot6d_to_rotmat(self, x6):
        n = x6.numel() // 6
        R = torch.empty(n, 3, 3, device=x6.device, dtype=torch.float32)
        if n == 0:
            return R
        check(self.lib.ehb_rot6d_to_rotmat(self._h, _dev_ptr(x6), n, _dev_ptr(R), _stream()))
        return R

    def rotmat_to_angle_axis(self, R):
        n = R.numel() // 9
        aa = torch.empty(n, 3, device=R.device, dtype=torch.float32)
        if n:
            check(self.lib.ehb_rotmat_to_angle_axis(self._h, _dev_ptr(R), n, _dev_ptr(aa), _stream()))
        return aa

    def cond_inputs(self, kp2d, scene_feat, transl_feat, img_feat, fx, box_center, box_size, cam_cx, cam_cy, flags,
                    openpose_to_smpl, fx_norm_coeff):
        """EgoHMR.forward's step-invariant glue in one launch (include/egohmr_b200.h::ehb_cond_inputs).
        flags = (with_focal_length, with_bbox_info, with_cam_center) -> vis uint8 [n,24], rest [n,R], ctx_full [n,img+R]."""
        n, sf, tf, img_dim = kp2d.shape[0], scene_feat.shape[1], transl_feat.shape[1], img_feat.shape[1]
        rdim = sf + tf + (1 if flags[0] else 0) + (3 if flags[1] else 0) + (2 if flags[2] else 0)
        dev = kp2d.device
        vis = torch.empty(n, 24, device=dev, dtype=torch.uint8)
        rest = torch.empty(n, rdim, device=dev, dtype=torch.float32)
        full = torch.empty(n, img_dim + rdim, device=dev, dtype=torch.float32)
        o2s = np.ascontiguousarray(openpose_to_smpl, dtype=np.int32)
        opt = lambda t: _dev_ptr(t, allow_none=True)
        check(self.lib.ehb_cond_inputs(self._h, _dev_ptr(kp2d), _dev_ptr(scene_feat), _dev_ptr(transl_feat), _dev_ptr(img_feat),
                                       opt(fx), opt(box_center), opt(box_size), opt(cam_cx), opt(cam_cy), n, sf, tf, img_dim,
                                       int(bool(flags[0])), int(bool(flags[1])), int(bool(flags[2])),
                                       o2s.ctypes.data_as(_lib.c_int32_p), float(fx_norm_coeff), _dev_ptr(vis, torch.uint8),
                                       _dev_ptr(rest), _dev_ptr(full), _stream()))
        return vis, rest, full

    def project_joints(self, joints, transl, fx, cam_cx, cam_cy, img_of_body, fx_norm_coeff, default_focal):
        """joints [B,J,3] + per-image translation / camera -> (kp3d_full [B,J,3], kp2d [B,J,2] normalised to the full frame,
        focal [B,2], centre [B,2]) in one launch (ehb_project_joints)."""
        B, J = joints.shape[0], joints.shape[1]
        dev = joints.device
        full = torch.empty(B, J, 3, device=dev, dtype=torch.float32)
        kp2d = torch.empty(B, J, 2, device=dev, dtype=torch.float32)
        focal = torch.empty(B, 2, device=dev, dtype=torch.float32)
        center = torch.empty(B, 2, device=dev, dtype=torch.float32)
        opt = lambda t: _dev_ptr(t, allow_none=True)
        check(self.lib.ehb_project_joints(self._h, _dev_ptr(joints), _dev_ptr(transl), opt(fx), opt(cam_cx), opt(cam_cy),
                                          _dev_ptr(img_of_body, torch.int32, allow_none=True), B, J, float(fx_norm_coeff),
                                          float(default_focal), _dev_ptr(full), _dev_ptr(kp2d), _dev_ptr(focal),
                                          _dev_ptr(center), _stream()))
        return full, kp2d, focal, center

    def scene_crop(self, verts, scene, img_of_body=None):
        """Bounding-box crop of guide_coll / eval_coll (egohmr.py:550-554) for all bodies at once.
        verts [B,V,3], scene [n_clouds,N,3], img_of_body int32 [B] (device) or None -> mask bool [B,N], count int32 [B]."""
        B, V = verts.shape[0], verts.shape[1]
        N = scene.shape[1]
        mask = torch.empty(B, N, device=verts.device, dtype=torch.uint8)
        count = torch.empty(B, device=verts.device, dtype=torch.int32)
        check(self.lib.ehb_scene_crop(self._h, _dev_ptr(verts), B, V, _dev_ptr(scene), N,
                                      _dev_ptr(img_of_body, torch.int32, allow_none=True), _dev_ptr(mask, torch.uint8),
                                      _dev_ptr(count, torch.int32), None, _stream()))
        return mask.bool(), count

    def procrustes(self, S1, S2, mask=None):
        """Batched similarity-transform alignment (utils/pose_utils.py:11-105): S1, S2 [P,N,3] (mask [P,N,3] optional)
        -> (S1_hat [P,N,3], err [P,N])."""
        P, N = S1.shape[0], S1.shape[1]
        hat = torch.empty(P, N, 3, device=S1.device, dtype=torch.float32)
        err = torch.empty(P, N, device=S1.device, dtype=torch.float32)
        if P:
            check(self.lib.ehb_procrustes_align(self._h, _dev_ptr(S1), _dev_ptr(S2), _dev_ptr(mask, allow_none=True), P, N,
                                                _dev_ptr(hat), _dev_ptr(err), _stream()))
        return hat, err

    def eval_metrics(self, pred_joints, pred_verts, transl, gt_joints, gt_verts, focal, cam_cx, cam_cy):
        """test_egohmr.py:373-494 for one batch (see include/egohmr_b200.h::ehb_eval_metrics).
        -> joint_vis bool [bs,J], vert_vis bool [bs,V], errors [bs,S,9], diversity [bs,6]"""
        bs, S, J = pred_joints.shape[:3]
        V = pred_verts.shape[2]
        dev = pred_joints.device
        jv = torch.empty(bs, J, device=dev, dtype=torch.uint8)
        vv = torch.empty(bs, V, device=dev, dtype=torch.uint8)
        err = torch.empty(bs, S, 9, device=dev, dtype=torch.float32)
        div = torch.empty(bs, 6, device=dev, dtype=torch.float32)
        check(self.lib.ehb_eval_metrics(self._h, _dev_ptr(pred_joints), _dev_ptr(pred_verts), _dev_ptr(transl),
                                        _dev_ptr(gt_joints), _dev_ptr(gt_verts), _dev_ptr(focal), _dev_ptr(cam_cx),
                                        _dev_ptr(cam_cy), bs, S, J, V, _dev_ptr(jv, torch.uint8), _dev_ptr(vv, torch.uint8),
                                        _dev_ptr(err), _dev_ptr(div), _stream()))
        return jv.bool(), vv.bool(), err, div

    def nn_dist_sq(self, q, r, q_index=None, r_index=None, n_pairs=None):
        """Squared 1-NN distances: q [Nq_clouds,Pq,3] vs r [Nr_clouds,Pr,3], pair i = (q[q_index[i]], r[r_index[i]])
        (identity when the index is None) -> [n_pairs, Pq]."""
        if n_pairs is None:
            n_pairs = (q_index if q_index is not None else (r_index if r_index is not None else q)).shape[0]
        out = torch.empty(n_pairs, q.shape[1], device=q.device, dtype=torch.float32)
        if n_pairs:
            check(self.lib.ehb_nn_dist_sq(self._h, _dev_ptr(q), _dev_ptr(q_index, torch.int32, allow_none=True), q.shape[1],
                                          _dev_ptr(r), _dev_ptr(r_index, torch.int32, allow_none=True), r.shape[1], n_pairs,
                                          _dev_ptr(out), _stream()))
        return out

    def smpl_backward(self, x_t, betas, g_verts=None, g_joints=None, g_aa=None):
        """dL/dx_t [B,144] for the bodies of set_bodies (guide_coll's autograd.grad, egohmr.py:562)."""
        grad = torch.empty_like(x_t)
        check(self.lib.ehb_smpl_backward(self._h, _dev_ptr(x_t), _dev_ptr(betas), _dev_ptr(g_verts, allow_none=True),
                                         _dev_ptr(g_joints, allow_none=True), _dev_ptr(g_aa, allow_none=True),
                                         _dev_ptr(grad), _stream()))
        return grad

    # ------------------------------------------------------------------ diagnostics
    def launch_count(self):
        return int(self.lib.ehb_launch_count(self._h))

    def alloc_epoch(self):
        return int(self.lib.ehb_alloc_epoch())

    def set_gemm_mode(self, mode):
        check(self.lib.ehb_debug_set_gemm_mode(self._h, int(mode)))

    def set_resnet_mode(self, implicit_gemm):
        check(self.lib.ehb_debug_set_resnet_mode(self._h, 1 if implicit_gemm else 0))

    def set_pdl(self, on):
        check(self.lib.ehb_debug_set_pdl(self._h, 1 if on else 0))

    def set_k1_fused(self, on):
        """Hidden layers of a step as one persistent launch (True; same bits, slower) or one launch per layer (False, default)."""
        check(self.lib.ehb_debug_set_k1_fused(self._h, 1 if on else 0))

    def set_input_mode(self, umma):
        """K2's joint mix on tcgen05 (True) or the fp32 FFMA kernel (False, default; measured equal)."""
        check(self.lib.ehb_debug_set_input_mode(self._h, 1 if umma else 0))

    def set_conv_kc(self, kc):
        """k-blocks (x64 operand columns) per tensor-memory accumulation chunk of the ResNet convolution GEMMs (0 = whole K)."""
        check(self.lib.ehb_debug_set_conv_kc(self._h, int(kc)))

    def debug_gemm_hl(self, a, w, a_scale=1.0, w_scale=1.0, kc=0):
        """out[m,n] = a[m,k] . w[n,k]^T through the convolution GEMM primitive (numpy in / out; unit tests of its numerics)."""
        a, w = f32(a), f32(w)
        m, k = a.shape
        n = w.shape[0]
        assert w.shape[1] == k
        out = np.empty((m, n), np.float32)
        check(self.lib.ehb_debug_gemm_hl(self._h, fptr(a), fptr(w), m, n, k, float(a_scale), float(w_scale), int(kc),
                                         fptr(out)))
        return out

    def check_overflow(self):
        return bool(self.lib.ehb_check_overflow(self._h, _stream()))

    def overflow_flag_async(self, host_flag):
        """Enqueue a copy of the fp16-overflow flag into `host_flag` (pinned int32 tensor) on the current stream."""
        assert host_flag.is_pinned() and host_flag.dtype == torch.int32
        check(self.lib.ehb_overflow_flag_async(self._h, C.c_void_p(host_flag.data_ptr()), _stream()))

    def time_stage(self, stage, step, x_t, iters):
        """Mean device ms of one stage of a reverse step (0 = K2 input layer, 1..8 = K1, 9 = K3 output + update)."""
        ms = C.c_float()
        x_prev, x0 = torch.empty_like(x_t), torch.empty_like(x_t)
        check(self.lib.ehb_time_stage(self._h, int(stage), int(step), _dev_ptr(x_t), _dev_ptr(x_prev), _dev_ptr(x0),
                                      int(iters), C.byref(ms), _stream()))
        return float(ms.value)

    def time_hidden_layer(self, layer, iters):
        ms = C.c_float()
        check(self.lib.ehb_time_hidden_layer(self._h, int(layer), int(iters), C.byref(ms), _stream()))
        return float(ms.value)
