"""Shim of timm.layers.conv_bn_act (test infrastructure, see ../__init__.py)."""
import torch
from torch import nn
import torch.nn.functional as F


class BatchNormAct2d(nn.BatchNorm2d):
    """BatchNorm2d followed by an optional activation (timm keeps both in `.bn`)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, apply_act=True):
        super().__init__(num_features, eps=eps, momentum=momentum)
        self.drop = nn.Identity()
        self.act = nn.ReLU(inplace=True) if apply_act else nn.Identity()

    def forward(self, x):
        x = super().forward(x)
        x = self.drop(x)
        return self.act(x)


class ConvNormAct(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, groups=1, apply_act=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                              padding=kernel_size // 2, groups=groups, bias=False)
        self.bn = BatchNormAct2d(out_channels, apply_act=apply_act)

    @property
    def in_channels(self):
        return self.conv.in_channels

    @property
    def out_channels(self):
        return self.conv.out_channels

    def forward(self, x):
        return self.bn(self.conv(x))


ConvBnAct = ConvNormAct


class SEModule(nn.Module):
    def __init__(self, channels, rd_channels):
        super().__init__()
        self.fc1 = nn.Conv2d(channels, rd_channels, kernel_size=1, bias=True)
        self.bn = nn.Identity()
        self.act = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(rd_channels, channels, kernel_size=1, bias=True)
        self.gate = nn.Sigmoid()

    def forward(self, x):
        x_se = x.mean((2, 3), keepdim=True)
        x_se = self.fc1(x_se)
        x_se = self.act(self.bn(x_se))
        x_se = self.fc2(x_se)
        return x * self.gate(x_se)
