"""The Modulated-GCN denoiser and EgoHMR.forward's denoising section (numpy).  TEST INFRASTRUCTURE (oracle/__init__).

Written in the reference's own (unfolded, 3718-wide) form so that it checks the product path's algebraic
restructuring rather than sharing it.  `sd` is a dict of numpy arrays keyed by the reference's state_dict names.
"""
import numpy as np


def modulated_graph_conv(x, W, M, adj, adj2, bias):
    """ModulatedGraphConv.forward (models/egohmr/modulated_gcn/modulated_gcn_conv.py:38-50).  x: [B, 24, in]."""
    dt = x.dtype
    h0 = np.matmul(x, W[0].astype(dt))
    h1 = np.matmul(x, W[1].astype(dt))
    a = adj.astype(dt) + adj2.astype(dt)
    a = (a.T + a) / dt.type(2)
    E = np.eye(a.shape[0], dtype=dt)
    M = M.astype(dt)
    out = np.matmul(a * E, M * h0) + np.matmul(a * (1 - E), M * h1)
    return out + bias.astype(dt).reshape(1, 1, -1)


def batchnorm_eval(x, sd, name, eps=1e-5):
    """nn.BatchNorm1d in eval mode over the channel (last) axis (modulated_gcn.py:22-23 transposes around it)."""
    dt = x.dtype
    w, b = sd[name + ".weight"].astype(dt), sd[name + ".bias"].astype(dt)
    m, v = sd[name + ".running_mean"].astype(dt), sd[name + ".running_var"].astype(dt)
    return (x - m) / np.sqrt(v + dt.type(eps)) * w + b


def graph_conv(x, sd, name, adj):
    """_GraphConv.forward (modulated_gcn.py:21-28): gconv -> BN -> ReLU (dropout p=0 is the identity in eval)."""
    y = modulated_graph_conv(x, sd[name + ".gconv.W"], sd[name + ".gconv.M"], adj, sd[name + ".gconv.adj2"],
                             sd[name + ".gconv.bias"])
    return np.maximum(batchnorm_eval(y, sd, name + ".bn"), 0)


def non_local_block(x, sd, name, eps=1e-5):
    """NONLocalBlock2D(sub_sample=False).forward on [B, 24, C] (nets/non_local_embedded_gaussian.py:61-85; the
    [B, C, 1, 24] layout of modulated_gcn.py:104-109 only moves axes around)."""
    dt = x.dtype
    conv = lambda t, n: np.matmul(t, sd[f"{name}.{n}.weight"].astype(dt).reshape(sd[f"{name}.{n}.weight"].shape[0], -1).T) \
        + sd[f"{name}.{n}.bias"].astype(dt)
    g_x, theta, phi = conv(x, "g"), conv(x, "theta"), conv(x, "phi")
    f = np.matmul(theta, np.swapaxes(phi, 1, 2))                       # [B, 24, 24]
    f = np.exp(f - f.max(axis=-1, keepdims=True))
    f = f / f.sum(axis=-1, keepdims=True)                              # F.softmax(f, dim=-1)
    y = np.matmul(f, g_x)
    return batchnorm_eval(conv(y, "W.0"), sd, f"{name}.W.1", eps) + x


def modulated_gcn(x, sd, adj, n_blocks, prefix="diffusion_model"):
    """ModulatedGCN.forward (modulated_gcn.py:99-116); the non-local block runs iff its parameters are in `sd`
    (nonlocal_layer=True; the reference's drivers leave it off, egohmr.py:37, test_egohmr.py:112-118)."""
    out = graph_conv(x, sd, f"{prefix}.gconv_input.0", adj)
    for b in range(n_blocks):  # _ResGraphConv.forward (modulated_gcn.py:38-42)
        res = out
        out = graph_conv(out, sd, f"{prefix}.gconv_layers.{b}.gconv1", adj)
        out = graph_conv(out, sd, f"{prefix}.gconv_layers.{b}.gconv2", adj)
        out = res + out
    if f"{prefix}.non_local.g.weight" in sd:
        out = non_local_block(out, sd, f"{prefix}.non_local")
    g = f"{prefix}.gconv_output"
    return modulated_graph_conv(out, sd[g + ".W"], sd[g + ".M"], adj, sd[g + ".adj2"], sd[g + ".bias"])


def linear(x, sd, name):
    dt = x.dtype
    y = np.matmul(x, sd[name + ".weight"].astype(dt).T)
    if name + ".bias" in sd:
        y = y + sd[name + ".bias"].astype(dt)
    return y


def timestep_embedding(t_orig, sd, dtype):
    """TimestepEmbedder.forward (egohmr.py:642-643): time_embed(pe[t]) -> [B, 512] after the squeeze at :178."""
    pe = sd["embed_timestep.sequence_pos_encoder.pe"].astype(dtype)[np.asarray(t_orig)][:, 0, :]
    h = linear(pe, sd, "embed_timestep.time_embed.0")
    h = h / (1 + np.exp(-h))  # SiLU
    return linear(h, sd, "embed_timestep.time_embed.2")


def denoise(sd, adj, n_blocks, x_t, t_orig, img_feats, rest_feats, vis, diffuse_fuse=True, only_mask_img_cond=True):
    """EgoHMR.forward's x_t-dependent part (egohmr.py:178-179, 190-191, 220-257) -> pred_x_start [B, 144]
    (plus the raw conditioned / image-masked outputs).

    img_feats [B,2048]; rest_feats [B,646] = [scene | transl | cam]; vis [B,24] bool; t_orig: ORIGINAL timesteps [B]
    (what _WrappedModel passes after timestep_map, respace.py:124-129)."""
    dt = x_t.dtype
    B = x_t.shape[0]
    temb = np.repeat(timestep_embedding(t_orig, sd, dt)[:, None, :], 24, axis=1)  # :178-179
    img_j = np.repeat(img_feats.astype(dt)[:, None, :], 24, axis=1) * vis[:, :, None].astype(dt)  # :190-191
    rest_j = np.repeat(rest_feats.astype(dt)[:, None, :], 24, axis=1)  # :220-222
    cond = np.concatenate([img_j, rest_j], axis=-1)  # :223
    x_feat = linear(x_t.reshape(B, 24, 6), sd, "input_process.poseEmbedding")  # :232-234
    feat = np.concatenate([cond, x_feat, temb], axis=-1)  # :236
    out_c = modulated_gcn(feat, sd, adj, n_blocks)  # :237
    if not diffuse_fuse:
        return out_c.reshape(B, 144), out_c.reshape(B, 144), None
    cond_u = cond.copy()
    if only_mask_img_cond:
        cond_u[:, :, 0:2048] = 0  # mask_cond(force_mask=True, only_mask_img_cond=True) :150-156, :242-244
    else:
        cond_u[:] = 0             # mask_cond(force_mask=True, only_mask_img_cond=False) :157-158
    out_u = modulated_gcn(np.concatenate([cond_u, x_feat, temb], axis=-1), sd, adj, n_blocks)  # :245-246
    guidance_param = 0  # :248
    out = out_u + guidance_param * (out_c - out_u)  # :249
    vis6 = np.repeat(vis[:, :, None], 6, axis=2).reshape(B, 144)  # :251
    out = out.reshape(B, 144)
    oc = out_c.reshape(B, 144)
    out[vis6] = oc[vis6]  # :254
    return out, oc, out_u.reshape(B, 144)
