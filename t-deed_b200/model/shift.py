"""Drop-in for the reference's model/shift.py: inserts Gate-Shift modules in stages s3 / s4."""
import math

from torch import nn

from model.impl.gsm import _GSM
from model.impl.gsf import _GSF
from model import regnet


def make_temporal_shift(net, clip_len, mode='gsm'):
    """Wrap conv1 of every block of s3 and s4 (shift.py:46-59 of the reference, n_round = 1)."""
    if mode not in ('gsm', 'gsf'):
        raise NotImplementedError('Unsupported shift mode')
    if not isinstance(net, regnet.RegNet):
        raise NotImplementedError('Unsupported architecture')
    for stage in (net.s3, net.s4):
        blocks = list(stage.children())
        print('=> Processing stage with {} blocks residual'.format(len(blocks)))
        for b in blocks:
            b.conv1 = GatedShift(b.conv1, n_segment=clip_len, n_div=4, mode=mode)


class GatedShift(nn.Module):
    """Parameter container `gs` (gate-shift) + `net` (the wrapped 1x1 ConvBnAct); shift.py:64-93."""

    def __init__(self, net, n_segment, n_div, mode='gsm'):
        super().__init__()
        if isinstance(net, regnet.ConvBnAct):
            channels = net.conv.in_channels
        elif isinstance(net, nn.Conv2d):
            channels = net.in_channels
        else:
            raise NotImplementedError(type(net))
        self.fold_dim = math.ceil(channels // n_div / 4) * 4
        if mode == 'gsm':
            self.gs = _GSM(self.fold_dim, n_segment)
        elif mode == 'gsf':
            self.gs = _GSF(self.fold_dim, n_segment, 100)
        self.net = net
        self.n_segment = n_segment
        print('=> Using GSM/GSF, fold dim: {} / {}'.format(self.fold_dim, channels))
