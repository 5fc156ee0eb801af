"""The reference's PyTorch dataflow, restated in eager torch (device-agnostic).  TEST INFRASTRUCTURE (oracle/__init__).

Purpose: SURVEY.md 8d asks for the reference's own *PyTorch-CUDA* path timed on the same B200 (the denominator of the
">= 10x" target).  `/root/reference` cannot travel to the GPU box and cannot be pip-installed, so this module restates
what `test_egohmr.py:251-255` executes — one `val_losses` call per sample, every reverse step re-running ResNet-50,
ResPointNet, both 3718-wide GCN passes, SMPL and the projection, the sampler update as ~15 small tensor ops with
`_extract_into_tensor` uploads — with stock torch ops and torch's default numerics flags.  Nothing is hoisted, folded
or fused: it is the baseline, not the product.  Checked against the reference's golden vectors in
tests/test_torch_eager_golden.py; timed by `bench.py --impl reference-cuda` (never imported by the product path).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import encoders as np_encoders
from . import schedule as np_schedule


class EagerEgoHMR:
    """EgoHMR.forward (models/egohmr/egohmr.py:173-303) from a reference-keyed state dict of numpy arrays."""

    def __init__(self, sd, adj, n_blocks, smpl_model, mean, std, device="cpu", dtype=torch.float32,
                 pelvis_vis_loosen=True, scene_cano=True, fx_norm_coeff=1500.0):
        self.dev, self.dt = torch.device(device), dtype
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.p = {k: (t(v).to(dtype) if np.asarray(v).dtype.kind == "f" else t(v)) for k, v in sd.items()}
        self.adj = t(adj).to(dtype)
        self.n_blocks = n_blocks
        self.mean, self.std = t(mean).to(dtype), t(std).to(dtype)
        self.smpl = {k: t(smpl_model[k]).to(dtype) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
        self.parents = [int(x) for x in smpl_model["parents"]]
        self.extra = [int(x) for x in smpl_model["extra_vertex_ids"]]
        table = np_encoders.OPENPOSE_TO_SMPL_LOOSEN if pelvis_vis_loosen else np_encoders.OPENPOSE_TO_SMPL
        self.op2smpl = table
        self.scene_cano, self.fx_norm = scene_cano, fx_norm_coeff

    # ---- small pieces
    def lin(self, x, name):
        return F.linear(x, self.p[name + ".weight"], self.p.get(name + ".bias"))

    def bn(self, x, name):
        p = self.p
        return F.batch_norm(x, p[name + ".running_mean"], p[name + ".running_var"], p[name + ".weight"], p[name + ".bias"],
                            False, 0.0, 1e-5)

    def resnet50(self, img, pre="backbone"):  # models/resnet.py:139-150
        x = F.relu(self.bn(F.conv2d(img, self.p[pre + ".conv1.weight"], stride=2, padding=3), pre + ".bn1"))
        x = F.max_pool2d(x, 3, 2, 1)
        for li, (planes, blocks, stride) in enumerate(np_encoders.RESNET50_LAYERS, start=1):
            for bi in range(blocks):
                q = f"{pre}.layer{li}.{bi}"
                s = stride if bi == 0 else 1
                y = F.relu(self.bn(F.conv2d(x, self.p[q + ".conv1.weight"]), q + ".bn1"))
                y = F.relu(self.bn(F.conv2d(y, self.p[q + ".conv2.weight"], stride=s, padding=1), q + ".bn2"))
                y = self.bn(F.conv2d(y, self.p[q + ".conv3.weight"]), q + ".bn3")
                if bi == 0:
                    x = self.bn(F.conv2d(x, self.p[q + ".downsample.0.weight"], stride=s), q + ".downsample.1")
                x = F.relu(y + x)
        return x.mean(dim=(2, 3))

    def block_fc(self, x, name):  # models/respointnet.py:88-97
        net = self.lin(F.relu(x), name + ".fc_0")
        dx = self.lin(F.relu(net), name + ".fc_1")
        return (self.lin(x, name + ".shortcut") if (name + ".shortcut.weight") in self.p else x) + dx

    def pointnet(self, pts, pre="scene_enc"):  # models/respointnet.py:33-59
        net = self.block_fc(self.lin(pts, pre + ".fc_pos_0"), pre + ".block_0")
        for b in (1, 2, 3):
            pooled = net.max(dim=1, keepdim=True)[0].expand(net.size())
            net = self.block_fc(torch.cat([net, pooled], dim=2), f"{pre}.block_{b}")
        return self.lin(F.relu(net.max(dim=1)[0]), pre + ".fc_c")

    def gconv(self, x, name):  # modulated_gcn_conv.py:38-50
        p = self.p
        h0 = torch.matmul(x, p[name + ".W"][0])
        h1 = torch.matmul(x, p[name + ".W"][1])
        adj = self.adj + p[name + ".adj2"]
        adj = (adj.T + adj) / 2
        E = torch.eye(adj.size(0), dtype=x.dtype, device=x.device)
        out = torch.matmul(adj * E, p[name + ".M"] * h0) + torch.matmul(adj * (1 - E), p[name + ".M"] * h1)
        return out + p[name + ".bias"].view(1, 1, -1)

    def graph_conv(self, x, name):  # modulated_gcn.py:21-28
        y = self.gconv(x, name + ".gconv").transpose(1, 2)
        return F.relu(self.bn(y, name + ".bn").transpose(1, 2))

    def gcn(self, x, pre="diffusion_model"):  # modulated_gcn.py:99-116
        out = self.graph_conv(x, pre + ".gconv_input.0")
        for b in range(self.n_blocks):
            res = out
            out = self.graph_conv(out, f"{pre}.gconv_layers.{b}.gconv1")
            out = self.graph_conv(out, f"{pre}.gconv_layers.{b}.gconv2")
            out = res + out
        return self.gconv(out, pre + ".gconv_output")

    def rot6d(self, x):  # utils/geometry.py:47-66, 'diffusion' mode
        x = x.reshape(-1, 3, 2)
        a1, a2 = x[:, :, 0], x[:, :, 1]
        b1 = F.normalize(a1)
        b2 = F.normalize(a2 - torch.einsum("bi,bi->b", b1, a2).unsqueeze(-1) * b1)
        b3 = torch.cross(b1, b2, dim=1)
        return torch.stack((b1, b2, b3), dim=-1)

    def smpl_forward(self, R, betas):  # smplx lbs (oracle/smpl.py), pose2rot=False
        m, B = self.smpl, R.shape[0]
        v_shaped = m["v_template"][None] + torch.einsum("bl,mkl->bmk", betas, m["shapedirs"])
        J = torch.einsum("bik,ji->bjk", v_shaped, m["J_regressor"])
        pf = (R[:, 1:] - torch.eye(3, dtype=R.dtype, device=R.device)).reshape(B, -1)
        v_posed = torch.matmul(pf, m["posedirs"]).reshape(B, -1, 3) + v_shaped
        rel = J.clone()
        rel[:, 1:] = J[:, 1:] - J[:, self.parents[1:]]
        T = torch.zeros(B, 24, 4, 4, dtype=R.dtype, device=R.device)
        T[:, :, :3, :3], T[:, :, :3, 3], T[:, :, 3, 3] = R, rel, 1
        chain = [T[:, 0]]
        for i in range(1, 24):
            chain.append(torch.matmul(chain[self.parents[i]], T[:, i]))
        G = torch.stack(chain, dim=1)
        posed = G[:, :, :3, 3]
        A = G.clone()
        A[:, :, :3, 3] = G[:, :, :3, 3] - torch.einsum("bjrc,bjc->bjr", G[:, :, :3, :3], J)
        Tv = torch.matmul(m["lbs_weights"], A.reshape(B, 24, 16)).reshape(B, -1, 4, 4)
        vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=R.dtype, device=R.device)], dim=2)
        verts = torch.matmul(Tv, vh.unsqueeze(-1))[:, :, :3, 0]
        return verts, torch.cat([posed, verts[:, self.extra]], dim=1)

    # ---- EgoHMR.forward
    def forward(self, batch, timesteps):
        p, dt = self.p, self.dt
        bs = batch["img"].shape[0]
        pe = p["embed_timestep.sequence_pos_encoder.pe"][timesteps]                       # [bs,1,512]
        temb = self.lin(F.silu(self.lin(pe, "embed_timestep.time_embed.0")), "embed_timestep.time_embed.2")
        temb = temb.permute(1, 0, 2).squeeze(0).unsqueeze(1).repeat(1, 24, 1)               # :178-179
        img_feats = self.resnet50(batch["img"].to(dt))                                     # :183
        vis_op = batch["orig_keypoints_2d"][:, :, -1] > 0
        vis_op[:, 8] = True
        vis = vis_op[:, self.op2smpl]
        img_j = img_feats.unsqueeze(1).repeat(1, 24, 1) * vis.unsqueeze(-1).repeat(1, 1, img_feats.shape[-1])
        fx = batch["fx"].to(dt)
        ofx = fx * self.fx_norm
        cam = [torch.stack([batch["cam_cx"].to(dt) / ofx, batch["cam_cy"].to(dt) / ofx], dim=-1),
               torch.stack([batch["box_center"][:, 0].to(dt) / ofx, batch["box_center"][:, 1].to(dt) / ofx,
                            batch["box_size"].to(dt) / ofx], dim=-1), fx.unsqueeze(1)]        # :195-205
        transl = batch["smpl_params"]["transl"].to(dt)
        pts = batch["scene_pcd_verts_full"].to(dt)
        if self.scene_cano:
            pts = pts - transl.unsqueeze(1)
        scene = self.pointnet(pts)                                                          # :214
        tr = self.lin(F.relu(self.lin(transl, "transl_enc.layers.0")), "transl_enc.layers.2")
        rest = torch.cat([scene, tr] + cam, dim=1)
        cond = torch.cat([img_j, rest.unsqueeze(1).repeat(1, 24, 1)], dim=-1)               # :220-223
        x_feat = self.lin(batch["x_t"].reshape(bs, 24, -1), "input_process.poseEmbedding")
        out_c = self.gcn(torch.cat([cond, x_feat, temb], dim=-1))                           # :236-237
        mask = torch.ones_like(cond)
        mask[:, :, 0:2048] = 0
        out_u = self.gcn(torch.cat([cond * mask, x_feat, temb], dim=-1))                    # :242-246
        out = out_u + 0 * (out_c - out_u)
        vis6 = vis.unsqueeze(-1).repeat(1, 1, 6).reshape(bs, -1)
        out = out.reshape(bs, -1)
        oc = out_c.clone().reshape(bs, -1)
        out[vis6] = oc[vis6]                                                                # :251-254
        res = {"pred_x_start": out}
        pose6d = out * self.std + self.mean
        R = self.rot6d(pose6d).view(bs, 24, 3, 3)
        feats_b = torch.cat([img_feats, scene, tr] + cam, dim=1)
        betas = self.lin(F.relu(self.lin(feats_b, "beta_layer.layers.0")), "beta_layer.layers.2") + p["beta_layer.init_betas"]
        res["pred_smpl_params"] = {"global_orient": R[:, [0]].clone(), "body_pose": R[:, 1:].clone(), "betas": betas.clone()}
        res["pred_pose_6d"] = pose6d
        verts, joints = self.smpl_forward(R.float().to(dt), betas)                          # :276
        res["pred_keypoints_3d"], res["pred_vertices"] = joints, verts
        focal = fx.unsqueeze(-1).repeat(1, 2) * self.fx_norm
        center = torch.cat([batch["cam_cx"].to(dt).unsqueeze(-1), batch["cam_cy"].to(dt).unsqueeze(-1)], dim=-1)
        res["pred_keypoints_3d_full"] = joints + transl.unsqueeze(1)
        rot = torch.eye(3, device=joints.device, dtype=dt).unsqueeze(0).expand(bs, -1, -1)  # geometry.py:78-116
        pc = torch.einsum("bij,bkj->bki", rot, joints) + transl.unsqueeze(1)
        pc = pc / pc[:, :, -1].unsqueeze(-1)
        K = torch.zeros(bs, 3, 3, dtype=dt, device=pc.device)
        K[:, 0, 0], K[:, 1, 1], K[:, 2, 2], K[:, :-1, -1] = focal[:, 0], focal[:, 1], 1.0, center
        kp = torch.einsum("bij,bkj->bki", K, pc)[:, :, :-1]
        kp[:, :, 0] = kp[:, :, 0] / 1920 - 0.5
        kp[:, :, 1] = kp[:, :, 1] / 1080 - 0.5
        res["pred_keypoints_2d_full"] = kp
        return res


def _extract(arr, t, shape):
    """_extract_into_tensor (gaussian_diffusion.py:784-797): a fresh pageable upload + gather + cast + expand per use."""
    r = torch.from_numpy(arr).to(device=t.device)[t].float()
    while len(r.shape) < len(shape):
        r = r[..., None]
    return r.expand(shape)


@torch.no_grad()
def val_losses(model, sch, batch, shape, mode="ddim", noise=None):
    """One chain: ddim_sample_loop / p_sample_loop (gaussian_diffusion.py:391-508, 618-718) -> last step's outputs.
    `noise` ([n_steps+1, B, 144], reference draw order) replaces torch's RNG when given (golden comparisons)."""
    dev = model.dev
    x = torch.randn(*shape, device=dev) if noise is None else noise[0].to(dev)
    x = x.to(model.dt)
    tmap = torch.tensor(sch.timestep_map, device=dev, dtype=torch.long)
    out = None
    for k, i in enumerate(range(sch.num_timesteps - 1, -1, -1)):
        t = torch.tensor([i] * shape[0], device=dev)
        batch["x_t"] = x
        out = model.forward(batch, tmap[t])                                                 # respace.py:124-129
        x0 = out["pred_x_start"]
        _ = _extract(sch.posterior_variance, t, x.shape)                                    # p_mean_variance :259-262
        _ = _extract(sch.posterior_log_variance_clipped, t, x.shape)
        mean = _extract(sch.posterior_mean_coef1, t, x.shape) * x0 + _extract(sch.posterior_mean_coef2, t, x.shape) * x
        eps_noise = (torch.randn_like(x) if noise is None else noise[1 + k].to(dev).to(x.dtype))
        nonzero = (t != 0).float().view(-1, *([1] * (len(x.shape) - 1)))
        if mode == "ddim":                                                                  # :511-556, eta = 0
            eps = (_extract(sch.sqrt_recip_alphas_cumprod, t, x.shape) * x - x0) / \
                _extract(sch.sqrt_recipm1_alphas_cumprod, t, x.shape)
            ab = _extract(sch.alphas_cumprod, t, x.shape)
            abp = _extract(sch.alphas_cumprod_prev, t, x.shape)
            sigma = 0.0 * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
            mean_pred = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps
            x = mean_pred + nonzero * sigma * eps_noise
        else:                                                                               # :298-337
            logvar = _extract(sch.posterior_log_variance_clipped, t, x.shape)
            x = mean + nonzero * torch.exp(0.5 * logvar) * eps_noise
    out["sample"] = x
    return out


def build(hid, n_blocks, T, respacing, device, seed=0, dtype=torch.float32):
    from egohmr_b200 import synth
    smpl_model = synth.make_smpl_model(seed)
    sd = synth.make_state_dict(seed, hid=hid, n_blocks=n_blocks, init_betas=smpl_model["init_betas"])
    mean, std = synth.body_rep_stats(seed)
    model = EagerEgoHMR(sd, synth.skeleton_adjacency(), n_blocks, smpl_model, mean, std, device, dtype)
    return model, np_schedule.Schedule(T, respacing)
