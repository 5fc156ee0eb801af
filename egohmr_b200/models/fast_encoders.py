"""Library (PyTorch / cuDNN / cuBLAS) inference forms of the two step-invariant feature providers.  NOT the
default: a sampling pass runs both encoders natively (K9 / K7 in the CUDA library); these forms are what
`EgoHMR.native_image_enc = False` / `native_scene_enc = False` select, kept as comparison points (bench.py times the pass with
the cuDNN TF32 image encoder as `eager_cudnn_tf32_enc`).  Arranged so that they do not dominate a pass either:

* ResNet-50: BatchNorm(eval) folded into the convolution weights/bias (exact in real arithmetic), channels_last so cuDNN
  does not transpose around every convolution;
* ResPointNet: `cat([net, pooled.expand])` followed by a Linear is split into a per-point GEMM on `net` plus a per-cloud
  row on `pooled` (models/respointnet.py:36-47: halves K of the fc_0 / shortcut GEMMs of blocks 1-3, exact algebra).

Both are rebuilt from the module's parameters whenever a state_dict is loaded, so reference checkpoints keep working.
"""
import torch
import torch.nn.functional as F


def _fold(conv, bn):
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = (conv.weight * s[:, None, None, None]).contiguous(memory_format=torch.channels_last)
    b = (bn.bias - bn.running_mean * s).contiguous()
    return w, b


class FoldedResNet50:
    def __init__(self, net, engine=None):
        self.engine = engine   # optional: the stem max-pool runs on the library's NHWC kernel (5x torch's)
        with torch.no_grad():
            self.stem = _fold(net.conv1, net.bn1)
            self.blocks = []
            for li in range(1, 5):
                for blk in getattr(net, f"layer{li}"):
                    ds = None if blk.downsample is None else (_fold(blk.downsample[0], blk.downsample[1]), blk.downsample[0].stride)
                    c3 = _fold(blk.conv3, blk.bn3)
                    # relu(conv3(y) + b3 + ds(x) + b_ds): the projection's bias rides on conv3's, so the projection is a
                    # bias-free convolution and no separate elementwise bias-add pass over the block output is needed
                    c3_gpu = c3 if ds is None else (c3[0], (c3[1] + ds[0][1]).contiguous())
                    self.blocks.append((_fold(blk.conv1, blk.bn1), _fold(blk.conv2, blk.bn2), blk.conv2.stride,
                                        c3, ds, c3_gpu))

    @torch.no_grad()
    def __call__(self, x):
        x = x.contiguous(memory_format=torch.channels_last)
        if x.is_cuda:
            # cuDNN's fused conv + bias (+ residual) + ReLU: one kernel per convolution instead of conv, bias-add, relu
            cr = lambda t, wb, stride, pad: torch.cudnn_convolution_relu(t, wb[0], wb[1], stride, pad, (1, 1), 1)
            x = cr(x, self.stem, (2, 2), (3, 3))
            x = self.engine.maxpool3x3s2(x) if self.engine is not None else F.max_pool2d(x, 3, 2, 1)
            for (c1, c2, s2, c3, ds, c3g) in self.blocks:
                y = cr(x, c1, (1, 1), (0, 0))
                y = cr(y, c2, tuple(s2), (1, 1))
                idt = x if ds is None else F.conv2d(x, ds[0][0], None, stride=ds[1])
                x = torch.cudnn_convolution_add_relu(y, c3g[0], idt, 1.0, c3g[1], (1, 1), (0, 0), (1, 1), 1)
            return x.mean(dim=(2, 3))
        x = F.relu_(F.conv2d(x, self.stem[0], self.stem[1], stride=2, padding=3))
        x = F.max_pool2d(x, 3, 2, 1)
        for (c1, c2, s2, c3, ds, _) in self.blocks:
            y = F.relu_(F.conv2d(x, c1[0], c1[1]))
            y = F.relu_(F.conv2d(y, c2[0], c2[1], stride=s2, padding=1))
            y = F.conv2d(y, c3[0], c3[1])
            idt = x if ds is None else F.conv2d(x, ds[0][0], ds[0][1], stride=ds[1])
            x = F.relu_(y.add_(idt))
        return x.mean(dim=(2, 3))


class SplitPointNet:
    def __init__(self, net):
        with torch.no_grad():
            self.pos = (net.fc_pos_0.weight.t().contiguous(), net.fc_pos_0.bias)
            self.blocks = []
            for i in range(4):
                b = getattr(net, f"block_{i}")
                self.blocks.append((b.fc_0.weight.t().contiguous(), b.fc_0.bias, b.fc_1.weight.t().contiguous(), b.fc_1.bias,
                                    b.shortcut.weight.t().contiguous()))
            self.out = (net.fc_c.weight.t().contiguous(), net.fc_c.bias)

    @torch.no_grad()
    def __call__(self, p):
        B, N, _ = p.shape
        net = torch.addmm(self.pos[1], p.reshape(B * N, 3), self.pos[0])             # [B*N, 512]
        w0, b0, w1, b1, ws = self.blocks[0]
        h = torch.addmm(b0, F.relu(net), w0)
        net = torch.addmm(b1, F.relu_(h), w1).add_(net @ ws)                          # [B*N, 256]
        for (w0, b0, w1, b1, ws) in self.blocks[1:]:
            H = net.shape[1]
            pooled = net.view(B, N, H).max(dim=1)[0]                                  # [B, 256]
            # fc_0(relu(cat[net, pooled])) = relu(net) W0[:H] + relu(pooled) W0[H:]   (per-cloud row broadcast)
            row0 = torch.addmm(b0, F.relu(pooled), w0[H:])
            h = (F.relu(net) @ w0[:H]).view(B, N, -1).add_(row0[:, None, :]).view(B * N, -1)
            rows = pooled @ ws[H:]
            sc = (net @ ws[:H]).view(B, N, -1).add_(rows[:, None, :]).view(B * N, -1)
            net = torch.addmm(b1, F.relu_(h), w1).add_(sc)
        pooled = net.view(B, N, -1).max(dim=1)[0]
        return torch.addmm(self.out[1], F.relu(pooled), self.out[0])
