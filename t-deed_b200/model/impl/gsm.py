"""Drop-in for the reference's model/impl/gsm.py: `_GSM` keeps the reference's parameter layout
(conv3D, bn) so checkpoints load strictly; its arithmetic (gsm.py:89-116) runs in
libtdeed_sm100's tdeed_gsf_fwd (mode GSM), fused into the following 1x1 conv as a K-segment."""
import torch
from torch import nn

from tdeed_b200 import _lib as L
from tdeed_b200 import ops


class _GateShiftBase(nn.Module):
    _mode = L.SHIFT_GSM

    def _prepared(self, device):
        eps = self.bn.eps
        scale = self.bn.weight.detach().float() / torch.sqrt(self.bn.running_var.float() + eps)
        shift = self.bn.bias.detach().float() - self.bn.running_mean.float() * scale
        p = dict(bn_scale=scale.to(device).contiguous(), bn_shift=shift.to(device).contiguous(),
                 w3d=self.conv3D.weight.detach().float().reshape(-1).to(device).contiguous(),
                 b3d=self.conv3D.bias.detach().float().to(device).contiguous())
        if self._mode == L.SHIFT_GSF:
            p['cc_w'] = torch.cat([self.channel_conv1.weight.detach().reshape(-1),
                                   self.channel_conv2.weight.detach().reshape(-1)]).float().to(device).contiguous()
            p['cc_b'] = torch.cat([self.channel_conv1.bias.detach(), self.channel_conv2.bias.detach()]).float().to(device).contiguous()
        return p

    def forward(self, x):
        """x: (B*T, F[+rest], h, w) NCHW CUDA tensor (eval mode) -> gate-shifted tensor of the same shape."""
        if not x.is_cuda:
            raise RuntimeError('tdeed_b200 has no CPU path: _GSM/_GSF.forward needs a CUDA tensor')
        if self.training:
            raise NotImplementedError('standalone gate-shift forward is inference-only; training goes through TDEEDModel')
        n, c, h, w = x.shape
        f = self.fPlane
        clips = n // self.num_segments
        xh = x[:, :f].permute(0, 2, 3, 1).contiguous().float()
        ws = torch.empty(ops.gsf_workspace_floats(clips, self.num_segments, h, w, f), dtype=torch.float32, device=x.device)
        ld = (f + 7) // 8 * 8
        out = torch.empty((n * h * w, ld), dtype=torch.float32, device=x.device)
        ops.gsf(xh, clips, self.num_segments, f, self._mode, self._prepared(x.device), ws, out)
        y = out[:, :f].reshape(n, h, w, f).permute(0, 3, 1, 2).to(x.dtype)
        return torch.cat([y, x[:, f:]], dim=1) if c > f else y


class _GSM(_GateShiftBase):
    def __init__(self, fPlane, num_segments=3):
        super().__init__()
        self.conv3D = nn.Conv3d(fPlane, 2, (3, 3, 3), stride=1, padding=(1, 1, 1), groups=2)
        nn.init.constant_(self.conv3D.weight, 0)
        nn.init.constant_(self.conv3D.bias, 0)
        self.fPlane = fPlane
        self.num_segments = num_segments
        self.bn = nn.BatchNorm3d(num_features=fPlane)
