"""RegNetY-200MF / 800MF parameter containers with timm's attribute names.

timm is not a dependency of this repo: the reference obtains its backbone from
`timm.create_model('regnety_002'|'regnety_008')` (model/model.py:37-46); this module declares the same
parameter tree (`stem.{conv,bn}`, `s{1..4}.b{k}.{conv1,conv2,conv3,downsample}.{conv,bn}`,
`se.{fc1,fc2}`, `head.fc`) so that reference checkpoints load strictly.  The modules are containers:
the arithmetic runs in libtdeed_sm100 (tdeed_stem_fwd / tdeed_gemm_fwd / tdeed_conv3x3g_fwd / tdeed_se_fwd).
"""
import math

from torch import nn

CFGS = {
    'regnety_002': dict(widths=[24, 56, 152, 368], depths=[1, 1, 4, 7], group_width=8),
    'regnety_008': dict(widths=[64, 128, 320, 768], depths=[1, 3, 8, 2], group_width=16),
}


class ConvBnAct(nn.Module):
    """conv (bias-free) + BatchNorm2d (+ ReLU): parameter container."""

    def __init__(self, cin, cout, k=1, stride=1, groups=1, apply_act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.apply_act = apply_act


class SEModule(nn.Module):
    def __init__(self, channels, rd_channels):
        super().__init__()
        self.fc1 = nn.Conv2d(channels, rd_channels, 1, bias=True)
        self.fc2 = nn.Conv2d(rd_channels, channels, 1, bias=True)


class Bottleneck(nn.Module):
    def __init__(self, cin, cout, stride, group_width):
        super().__init__()
        self.conv1 = ConvBnAct(cin, cout, 1)
        self.conv2 = ConvBnAct(cout, cout, 3, stride=stride, groups=cout // group_width)
        self.se = SEModule(cout, int(round(cin * 0.25)))
        self.conv3 = ConvBnAct(cout, cout, 1, apply_act=False)
        self.downsample = ConvBnAct(cin, cout, 1, stride=stride, apply_act=False) if (cin != cout or stride != 1) \
            else nn.Identity()


class RegStage(nn.Module):
    def __init__(self, depth, cin, cout, group_width):
        super().__init__()
        for i in range(depth):
            self.add_module('b%d' % (i + 1), Bottleneck(cin if i == 0 else cout, cout, 2 if i == 0 else 1, group_width))


class Head(nn.Module):
    def __init__(self, in_features, num_classes=1000):
        super().__init__()
        self.fc = nn.Linear(in_features, num_classes)


class RegNet(nn.Module):
    def __init__(self, name):
        super().__init__()
        cfg = CFGS[name]
        self.stem = ConvBnAct(3, 32, 3, stride=2)
        prev = 32
        for i, (w, d) in enumerate(zip(cfg['widths'], cfg['depths'])):
            self.add_module('s%d' % (i + 1), RegStage(d, prev, w, cfg['group_width']))
            prev = w
        self.final_conv = nn.Identity()
        self.num_features = prev
        self.head = Head(prev)
        # timm's init: conv ~ N(0, sqrt(2/fan_out)), BN (1, 0), zero-init of every block's last BN gamma
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, Bottleneck):
                nn.init.zeros_(m.conv3.bn.weight)


PRETRAINED_ENV = 'TDEED_REGNET_WEIGHTS'


def create_model(name, pretrained=False, pretrained_path=None):
    """Stand-in for timm.create_model (model/model.py:37-46 of the reference calls it with pretrained=True, i.e. it
    starts from timm's ImageNet weights).  There is no network here, so the weights come from a LOCAL file:
    `pretrained_path`, or `$TDEED_REGNET_WEIGHTS` — a directory holding `<name>.pth` / `<name>.pt` / `<name>.bin`, or that
    file itself — containing timm's state_dict for `name` (keys `stem.conv.weight`, `s1.b1.conv1.conv.weight`, ...,
    `head.fc.*`); it is loaded STRICTLY.  With pretrained=True and no file this warns loudly and returns the random
    initialisation: fine for loading a full T-DEED checkpoint afterwards (`TDEEDModel.load`, what evaluation does), NOT
    equivalent to the reference for training from scratch."""
    model = RegNet(name)
    if not pretrained:
        return model
    import os
    import warnings
    path = pretrained_path or os.environ.get(PRETRAINED_ENV)
    if path and os.path.isdir(path):
        cands = [os.path.join(path, name + ext) for ext in ('.pth', '.pt', '.bin')]
        path = next((c for c in cands if os.path.isfile(c)), None)
        if path is None:
            raise FileNotFoundError('%s is set but holds none of %s' % (PRETRAINED_ENV, [os.path.basename(c) for c in cands]))
    if not path:
        msg = ('tdeed_b200: create_model(%r, pretrained=True) but no local weights: the backbone is RANDOMLY initialised (timm would '
               'have downloaded ImageNet weights).  Set %s=<dir or file with timm\'s %s state_dict> to match the reference when '
               'training from scratch; loading a T-DEED checkpoint afterwards overwrites the backbone anyway.' % (name, PRETRAINED_ENV, name))
        warnings.warn(msg, RuntimeWarning, stacklevel=2)
        print('WARNING: ' + msg)
        return model
    import torch
    sd = torch.load(path, map_location='cpu')
    if isinstance(sd, dict) and 'state_dict' in sd and not any(k.startswith('stem.') for k in sd):
        sd = sd['state_dict']
    model.load_state_dict(sd, strict=True)
    return model
