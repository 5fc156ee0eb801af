"""ResNet-50 image encoder: a step-invariant *feature provider* for the sampling hot path.

It is evaluated once per image, outside the diffusion loop (the reference re-runs it on every step, egohmr.py:183).
This nn.Module only HOLDS the parameters (and is the fp32 / fp64 comparison form in the tests): on a sampling pass the
encoder runs as tcgen05 convolution GEMMs in the CUDA library (`ehb_resnet_forward`, csrc/conv_umma.cu), unless
`EgoHMR.native_image_enc = False` selects the cuDNN inference form of models/fast_encoders.py.  Parameter names match the reference's `backbone.*` state_dict keys
(models/resnet.py:100-150: torchvision-style v1.5 bottlenecks, global average pool, no fc)."""
import torch.nn as nn

_STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))


class _Bottleneck(nn.Module):
    def __init__(self, cin, planes, stride, project):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if project:
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes * 4))

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + (x if self.downsample is None else self.downsample(x)))


class ResNet50Features(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin = 64
        for i, (planes, blocks, stride) in enumerate(_STAGES, start=1):
            layers = []
            for b in range(blocks):
                layers.append(_Bottleneck(cin, planes, stride if b == 0 else 1, b == 0))
                cin = planes * 4
            setattr(self, f"layer{i}", nn.Sequential(*layers))

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return x.mean(dim=(2, 3))
