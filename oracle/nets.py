"""Oracle: DispResNet / PoseResNet in plain PyTorch (restates reference models/*.py and the
torchvision ResNet blocks they instantiate).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Module attribute names reproduce the reference's `state_dict` keys exactly
(`encoder.encoder.layer1.0.conv1.weight`, `decoder.decoder.3.conv.conv.bias`,
`decoder.net.0.weight`, ...) so that weights interchange.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

STAGE_BLOCKS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


class _Basic(nn.Module):
    """Two 3x3 conv+BN with identity/projection shortcut (torchvision models/resnet.py:59-105)."""
    expansion = 1

    def __init__(self, cin, width, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.downsample = None
        if stride != 1 or cin != width:
            self.downsample = nn.Sequential(nn.Conv2d(cin, width, 1, stride, bias=False), nn.BatchNorm2d(width))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        sc = x if self.downsample is None else self.downsample(x)
        return F.relu(y + sc)


class _Bottleneck(nn.Module):
    """1x1 -> 3x3 (carries the stride) -> 1x1(x4) with shortcut (torchvision models/resnet.py:108-163)."""
    expansion = 4

    def __init__(self, cin, width, stride):
        super().__init__()
        cout = width * 4
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        sc = x if self.downsample is None else self.downsample(x)
        return F.relu(y + sc)


class _ResNetTrunk(nn.Module):
    """conv1/bn1/maxpool/layer1..4 (+ the unused fc that the reference keeps in its state_dict).

    Initialisation follows torchvision's ResNet.__init__ / reference resnet_encoder.py:34-39:
    kaiming-normal (fan_out, relu) conv weights, BN weight 1 / bias 0.
    """

    def __init__(self, num_layers, in_ch):
        super().__init__()
        block = _Basic if num_layers < 50 else _Bottleneck
        self.conv1 = nn.Conv2d(in_ch, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for i, (n, width) in enumerate(zip(STAGE_BLOCKS[num_layers], (64, 128, 256, 512))):
            blocks = []
            for j in range(n):
                blocks.append(block(cin, width, 2 if (j == 0 and i > 0) else 1))
                cin = width * block.expansion
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
        self.fc = nn.Linear(cin, 1000)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class ResnetEncoder(nn.Module):
    """Five feature maps at strides 2..32 (reference resnet_encoder.py:62-97)."""

    def __init__(self, num_layers, pretrained=False, num_input_images=1):
        super().__init__()
        if num_layers not in STAGE_BLOCKS:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        if pretrained:
            raise RuntimeError("oracle: ImageNet weights are not available offline; use pretrained=False")
        self.num_ch_enc = [64, 64, 128, 256, 512] if num_layers <= 34 else [64, 256, 512, 1024, 2048]
        self.encoder = _ResNetTrunk(num_layers, 3 * num_input_images)

    def forward(self, x):
        e = self.encoder
        f0 = F.relu(e.bn1(e.conv1(x)))
        f1 = e.layer1(F.max_pool2d(f0, 3, 2, 1))
        f2 = e.layer2(f1)
        f3 = e.layer3(f2)
        f4 = e.layer4(f3)
        return [f0, f1, f2, f3, f4]


class _ReflConv(nn.Module):
    """ReflectionPad2d(1) + 3x3 conv with bias (reference DispResNet.py:27-42)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(int(cin), int(cout), 3)

    def forward(self, x):
        return self.conv(F.pad(x, (1, 1, 1, 1), mode="reflect"))


class _ReflConvELU(nn.Module):
    """_ReflConv followed by ELU (reference DispResNet.py:13-25)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _ReflConv(cin, cout)

    def forward(self, x):
        return F.elu(self.conv(x))


class DepthDecoder(nn.Module):
    """monodepth2-style decoder (reference DispResNet.py:49-101): ModuleList order is
    upconv(4,0),(4,1),(3,0)...(0,1), then dispconv 0..3; disparity = 10*sigmoid(.)+0.01."""

    def __init__(self, num_ch_enc):
        super().__init__()
        dec = [16, 32, 64, 128, 256]
        mods, self._idx = [], {}
        for i in range(4, -1, -1):
            cin = num_ch_enc[-1] if i == 4 else dec[i + 1]
            self._idx[("up", i, 0)] = len(mods)
            mods.append(_ReflConvELU(cin, dec[i]))
            cin = dec[i] + (num_ch_enc[i - 1] if i > 0 else 0)
            self._idx[("up", i, 1)] = len(mods)
            mods.append(_ReflConvELU(cin, dec[i]))
        for s in range(4):
            self._idx[("disp", s)] = len(mods)
            mods.append(_ReflConv(dec[s], 1))
        self.decoder = nn.ModuleList(mods)

    def forward(self, feats):
        outs = {}
        x = feats[-1]
        for i in range(4, -1, -1):
            x = self.decoder[self._idx[("up", i, 0)]](x)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            if i > 0:
                x = torch.cat([x, feats[i - 1]], 1)
            x = self.decoder[self._idx[("up", i, 1)]](x)
            if i < 4:
                outs[i] = 10 * torch.sigmoid(self.decoder[self._idx[("disp", i)]](x)) + 0.01
        return [outs[s] for s in range(4)]


class DispResNet(nn.Module):
    """reference DispResNet.py:104-121: list of 4 disparities in train mode, scale 0 only in eval."""

    def __init__(self, num_layers=18, pretrained=False):
        super().__init__()
        self.encoder = ResnetEncoder(num_layers, pretrained, 1)
        self.decoder = DepthDecoder(self.encoder.num_ch_enc)

    def init_weights(self):
        pass

    def forward(self, x):
        outs = self.decoder(self.encoder(x))
        return outs if self.training else outs[0]


class PoseDecoder(nn.Module):
    """reference PoseResNet.py:14-51: 1x1 squeeze, two 3x3, 1x1 -> 6, spatial mean, x0.01."""

    def __init__(self, num_ch_enc):
        super().__init__()
        self.net = nn.ModuleList([nn.Conv2d(num_ch_enc[-1], 256, 1), nn.Conv2d(256, 256, 3, 1, 1),
                                  nn.Conv2d(256, 256, 3, 1, 1), nn.Conv2d(256, 6, 1)])

    def forward(self, feat):
        x = F.relu(self.net[0](feat))
        x = F.relu(self.net[1](x))
        x = F.relu(self.net[2](x))
        x = self.net[3](x)
        return 0.01 * x.mean(3).mean(2).view(-1, 6)


class PoseResNet(nn.Module):
    """reference PoseResNet.py:54-68: channel-concatenated image pair -> [B,6]."""

    def __init__(self, num_layers=18, pretrained=False):
        super().__init__()
        self.encoder = ResnetEncoder(num_layers, pretrained, 2)
        self.decoder = PoseDecoder(self.encoder.num_ch_enc)

    def init_weights(self):
        pass

    def forward(self, img1, img2):
        return self.decoder(self.encoder(torch.cat([img1, img2], 1))[-1])
