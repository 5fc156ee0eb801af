"""ResNet-PointNet scene encoder: step-invariant feature provider, evaluated once per image.
Parameter names match the reference's `scene_enc.*` keys (models/respointnet.py:13-27, 76-86)."""
import torch
import torch.nn as nn


class _BlockFC(nn.Module):
    def __init__(self, size_in, size_out, size_h):
        super().__init__()
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        self.shortcut = None if size_in == size_out else nn.Linear(size_in, size_out, bias=False)

    def forward(self, x):
        dx = self.fc_1(torch.relu(self.fc_0(torch.relu(x))))
        return (x if self.shortcut is None else self.shortcut(x)) + dx


class ResnetPointnet(nn.Module):
    def __init__(self, out_dim, hidden_dim):
        super().__init__()
        self.fc_pos_0 = nn.Linear(3, 2 * hidden_dim)
        for i in range(4):
            setattr(self, f"block_{i}", _BlockFC(2 * hidden_dim, hidden_dim, hidden_dim))
        self.fc_c = nn.Linear(hidden_dim, out_dim)

    def forward(self, p):
        net = self.block_0(self.fc_pos_0(p))
        for blk in (self.block_1, self.block_2, self.block_3):
            pooled = net.max(dim=1, keepdim=True)[0].expand(net.size())
            net = blk(torch.cat([net, pooled], dim=2))
        return self.fc_c(torch.relu(net.max(dim=1)[0]))
