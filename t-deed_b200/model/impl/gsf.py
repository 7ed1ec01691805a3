"""Drop-in for the reference's model/impl/gsf.py: `_GSF` keeps the reference's parameter layout
(conv3D, bn, channel_conv1/2); its arithmetic (gsf.py:38-93) runs in tdeed_gsf_fwd (mode GSF)."""
from torch import nn

from tdeed_b200 import _lib as L
from .gsm import _GateShiftBase


class _GSF(_GateShiftBase):
    _mode = L.SHIFT_GSF

    def __init__(self, fPlane, num_segments=8, gsf_ch_ratio=100):
        super().__init__()
        fPlane_temp = int(fPlane * gsf_ch_ratio / 100)
        if fPlane_temp % 2 != 0:
            fPlane_temp += 1
        self.fPlane = fPlane_temp
        self.conv3D = nn.Conv3d(self.fPlane, 2, (3, 3, 3), stride=1, padding=(1, 1, 1), groups=2)
        self.num_segments = num_segments
        self.bn = nn.BatchNorm3d(num_features=self.fPlane)
        self.channel_conv1 = nn.Conv2d(2, 1, (3, 3), padding=(3 // 2, 3 // 2))
        self.channel_conv2 = nn.Conv2d(2, 1, (3, 3), padding=(3 // 2, 3 // 2))
