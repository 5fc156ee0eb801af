"""Step-invariant feature providers of EgoHMR.forward (numpy; convolutions borrow torch's CPU conv2d as a numeric
library).  TEST INFRASTRUCTURE (see oracle/__init__).  These are not on the accelerated hot path but
the reference recomputes them on every diffusion step, so the CPU baseline needs them."""
import numpy as np

from .gcn import linear

OPENPOSE_TO_SMPL = [8, 12, 9, 8, 13, 10, 8, 14, 11, 8, 14, 11, 0, 5, 2, 0, 5, 2, 6, 3, 7, 4, 7, 4]          # egohmr.py:111
OPENPOSE_TO_SMPL_LOOSEN = [8, 13, 10, 8, 13, 10, 8, 14, 11, 8, 14, 11, 1, 5, 2, 0, 5, 2, 6, 3, 7, 4, 7, 4]   # egohmr.py:114
RESNET50_LAYERS = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]


def vis_mask_smpl(orig_keypoints_2d, pelvis_vis_loosen=True):
    """egohmr.py:186-189."""
    vis = orig_keypoints_2d[:, :, -1] > 0
    vis = vis.copy()
    vis[:, 8] = True
    table = OPENPOSE_TO_SMPL_LOOSEN if pelvis_vis_loosen else OPENPOSE_TO_SMPL
    return vis[:, table]


def cam_feats(batch, dtype, fx_norm_coeff=1500.0):
    """egohmr.py:195-205 with with_focal_length = with_bbox_info = with_cam_center = True ->
    [cam_cx/fx', cam_cy/fx', box_cx/fx', box_cy/fx', box_size/fx', fx]."""
    fx = batch["fx"].astype(dtype)
    orig_fx = fx * dtype(fx_norm_coeff)
    bc = batch["box_center"].astype(dtype)
    return np.stack([batch["cam_cx"].astype(dtype) / orig_fx, batch["cam_cy"].astype(dtype) / orig_fx,
                     bc[:, 0] / orig_fx, bc[:, 1] / orig_fx, batch["box_size"].astype(dtype) / orig_fx, fx], axis=-1)


def _bn2d(x, sd, name, eps=1e-5):
    dt = x.dtype
    w, b = sd[name + ".weight"].astype(dt), sd[name + ".bias"].astype(dt)
    m, v = sd[name + ".running_mean"].astype(dt), sd[name + ".running_var"].astype(dt)
    sh = (1, -1, 1, 1)
    return (x - m.reshape(sh)) / np.sqrt(v.reshape(sh) + dt.type(eps)) * w.reshape(sh) + b.reshape(sh)


def _conv(x, w, stride=1, padding=0):
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        y = F.conv2d(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(w.astype(x.dtype))),
                     stride=stride, padding=padding)
    return y.numpy()


def _maxpool3x3s2(x):
    import torch
    import torch.nn.functional as F
    return F.max_pool2d(torch.from_numpy(x), kernel_size=3, stride=2, padding=1).numpy()


def resnet50_features(sd, img, prefix="backbone"):
    """models/resnet.py:139-150 (ResNet-50 v1.5 bottlenecks :60-97) -> [B, 2048]."""
    x = _conv(img, sd[prefix + ".conv1.weight"], 2, 3)
    x = np.maximum(_bn2d(x, sd, prefix + ".bn1"), 0)
    x = _maxpool3x3s2(x)
    for li, (planes, blocks, stride) in enumerate(RESNET50_LAYERS, start=1):
        for bi in range(blocks):
            p = f"{prefix}.layer{li}.{bi}"
            s = stride if bi == 0 else 1
            out = np.maximum(_bn2d(_conv(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"), 0)
            out = np.maximum(_bn2d(_conv(out, sd[p + ".conv2.weight"], s, 1), sd, p + ".bn2"), 0)
            out = _bn2d(_conv(out, sd[p + ".conv3.weight"]), sd, p + ".bn3")
            if bi == 0:
                x = _bn2d(_conv(x, sd[p + ".downsample.0.weight"], s, 0), sd, p + ".downsample.1")
            x = np.maximum(out + x, 0)
    return x.mean(axis=(2, 3))


def _resblock_fc(x, sd, name):
    """ResnetBlockFC.forward (models/respointnet.py:88-97)."""
    net = linear(np.maximum(x, 0), sd, name + ".fc_0")
    dx = linear(np.maximum(net, 0), sd, name + ".fc_1")
    xs = linear(x, sd, name + ".shortcut") if (name + ".shortcut.weight") in sd else x
    return xs + dx


def respointnet(sd, p, prefix="scene_enc"):
    """ResnetPointnet.forward (models/respointnet.py:33-59).  p: [B, N, 3] -> [B, 512]."""
    net = linear(p, sd, prefix + ".fc_pos_0")
    net = _resblock_fc(net, sd, prefix + ".block_0")
    for b in (1, 2, 3):
        pooled = np.broadcast_to(net.max(axis=1, keepdims=True), net.shape)
        net = _resblock_fc(np.concatenate([net, pooled], axis=2), sd, f"{prefix}.block_{b}")
    net = net.max(axis=1)
    return linear(np.maximum(net, 0), sd, prefix + ".fc_c")


def transl_enc(sd, transl):
    """TranslEnc.forward (egohmr.py:685-691)."""
    return linear(np.maximum(linear(transl, sd, "transl_enc.layers.0"), 0), sd, "transl_enc.layers.2")


def beta_head(sd, feats):
    """FCHeadBeta.forward (egohmr.py:673-679), condition_on_pose=False."""
    h = np.maximum(linear(feats, sd, "beta_layer.layers.0"), 0)
    return linear(h, sd, "beta_layer.layers.2") + sd["beta_layer.init_betas"].astype(feats.dtype)


def conditioning(sd, batch, dtype=np.float32, scene_cano=True, pelvis_vis_loosen=True):
    """Everything in EgoHMR.forward that does not depend on x_t / t (egohmr.py:181-223, 262-265)."""
    img_feats = resnet50_features(sd, batch["img"].astype(dtype))
    vis = vis_mask_smpl(batch["orig_keypoints_2d"], pelvis_vis_loosen)
    transl = batch["smpl_params"]["transl"].astype(dtype)
    pts = batch["scene_pcd_verts_full"].astype(dtype)
    if scene_cano:
        pts = pts - transl[:, None, :]
    scene = respointnet(sd, pts)
    tr = transl_enc(sd, transl)
    cam = cam_feats(batch, np.dtype(dtype).type)
    rest = np.concatenate([scene, tr, cam], axis=1)
    betas = beta_head(sd, np.concatenate([img_feats, rest], axis=1))
    return {"img_feats": img_feats, "rest_feats": rest, "vis": vis, "betas": betas, "scene_pts": pts, "transl": transl}
