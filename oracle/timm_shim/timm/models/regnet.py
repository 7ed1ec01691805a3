"""Shim of timm.models.regnet for regnety_002 / regnety_008 (test infrastructure).

Architecture (pycls RegNetY, as shipped by timm 1.0.3):
  stem  ConvNormAct(3->32, k3, s2)
  s1..s4 stages of `Bottleneck` blocks b1..bD, first block of every stage stride 2
  Bottleneck: conv1 1x1(+BN+ReLU) -> conv2 3x3 grouped, stride s (+BN+ReLU) -> SE
              (rd = round(0.25 * block_in_chs)) -> conv3 1x1 (+BN) -> + shortcut -> ReLU
  shortcut  : ConvNormAct 1x1 stride s without activation when shape changes, else Identity
  final_conv Identity, head = global average pool + fc(1000)
  regnety_002: widths [24,56,152,368], depths [1,1,4,7], group width 8
  regnety_008: widths [64,128,320,768], depths [1,3,8,2], group width 16
"""
from torch import nn

from ..layers.conv_bn_act import ConvNormAct, SEModule

_CFGS = {
    'regnety_002': dict(widths=[24, 56, 152, 368], depths=[1, 1, 4, 7], group_size=8),
    'regnety_008': dict(widths=[64, 128, 320, 768], depths=[1, 3, 8, 2], group_size=16),
}


class Bottleneck(nn.Module):
    def __init__(self, in_chs, out_chs, stride, group_size, se_ratio=0.25):
        super().__init__()
        groups = out_chs // group_size
        self.conv1 = ConvNormAct(in_chs, out_chs, 1)
        self.conv2 = ConvNormAct(out_chs, out_chs, 3, stride=stride, groups=groups)
        self.se = SEModule(out_chs, rd_channels=int(round(in_chs * se_ratio)))
        self.conv3 = ConvNormAct(out_chs, out_chs, 1, apply_act=False)
        self.act3 = nn.ReLU(inplace=True)
        if in_chs != out_chs or stride != 1:
            self.downsample = ConvNormAct(in_chs, out_chs, 1, stride=stride, apply_act=False)
        else:
            self.downsample = nn.Identity()
        self.drop_path = nn.Identity()

    def forward(self, x):
        shortcut = x
        x = self.conv1(x)
        x = self.conv2(x)
        x = self.se(x)
        x = self.conv3(x)
        x = self.drop_path(x) + self.downsample(shortcut)
        return self.act3(x)


class RegStage(nn.Module):
    def __init__(self, depth, in_chs, out_chs, group_size):
        super().__init__()
        for i in range(depth):
            self.add_module('b{}'.format(i + 1),
                            Bottleneck(in_chs if i == 0 else out_chs, out_chs,
                                       stride=2 if i == 0 else 1, group_size=group_size))

    def forward(self, x):
        for block in self.children():
            x = block(x)
        return x


class ClassifierHead(nn.Module):
    def __init__(self, in_features, num_classes):
        super().__init__()
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.drop = nn.Identity()
        self.fc = nn.Linear(in_features, num_classes)
        self.flatten = nn.Flatten(1)

    def forward(self, x):
        x = self.flatten(self.global_pool(x))
        return self.fc(self.drop(x))


class RegNet(nn.Module):
    def __init__(self, widths, depths, group_size, num_classes=1000):
        super().__init__()
        self.stem = ConvNormAct(3, 32, 3, stride=2)
        prev = 32
        for i, (w, d) in enumerate(zip(widths, depths)):
            self.add_module('s{}'.format(i + 1), RegStage(d, prev, w, group_size))
            prev = w
        self.final_conv = nn.Identity()
        self.num_features = prev
        self.head = ClassifierHead(prev, num_classes)
        # timm: zero_init_last=True zeroes the last BN gamma of every block
        for m in self.modules():
            if isinstance(m, Bottleneck):
                nn.init.zeros_(m.conv3.bn.weight)

    def forward(self, x):
        x = self.stem(x)
        x = self.s1(x)
        x = self.s2(x)
        x = self.s3(x)
        x = self.s4(x)
        x = self.final_conv(x)
        return self.head(x)


def create_model(name, pretrained=False, **kwargs):
    """`pretrained` is ignored: there is no network and no weight file."""
    return RegNet(**_CFGS[name])
