"""Drop-in for the reference's model/modules.py (same public names, parameter names and call
signatures).  The nn.Modules are parameter containers whose forward methods launch the sm_100a
kernels of libtdeed_sm100 through tdeed_b200.ops; there is no PyTorch / CPU fallback.
"""
import abc
import math

import torch
import torch.nn as nn

from tdeed_b200 import _lib as L
from tdeed_b200 import ops
from tdeed_b200.engine import sgp_up_size


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError('%s: tdeed_b200 has no CPU path, got a %s tensor' % (what, t.device))


def _gemm_dtype():
    """bf16 tensor-core GEMMs under autocast (the reference trains/infers under fp16 autocast), exact fp32 otherwise."""
    return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32


class ABCModel:

    @abc.abstractmethod
    def get_optimizer(self, opt_args):
        raise NotImplementedError()

    @abc.abstractmethod
    def epoch(self, loader, **kwargs):
        raise NotImplementedError()

    @abc.abstractmethod
    def predict(self, seq):
        raise NotImplementedError()

    @abc.abstractmethod
    def state_dict(self):
        raise NotImplementedError()

    @abc.abstractmethod
    def load(self, state_dict):
        raise NotImplementedError()


class BaseRGBModel(ABCModel):

    def get_optimizer(self, opt_args):
        # bf16 autocast needs no loss scaling: a disabled GradScaler keeps `step()` call-compatible
        # with the reference (modules.py:37-39 returns GradScaler() iff device == 'cuda').
        # The optimizer is a real torch.optim.Optimizer (schedulers / state_dict work) whose step is ONE fused
        # tdeed_adamw_step launch over the flat parameter buffer (tdeed_b200/optim.py).
        from tdeed_b200.optim import FusedAdamW
        flat = self._model.flat_params() if hasattr(self._model, 'flat_params') else None
        if hasattr(self._model, 'sync_replicas'):
            self._model.sync_replicas()        # data parallel: every replica starts from rank 0's weights and BN buffers
        return FusedAdamW(self._get_params(), flat=flat, **opt_args), \
            torch.amp.GradScaler('cuda', enabled=False) if self.device == 'cuda' else None

    """ Assume there is a self._model """

    def _get_params(self):
        return list(self._model.parameters())

    def state_dict(self):
        if isinstance(self._model, nn.DataParallel):
            return self._model.module.state_dict()
        return self._model.state_dict()

    def load(self, state_dict):
        if isinstance(self._model, nn.DataParallel):
            self._model.module.load_state_dict(state_dict)
        else:
            self._model.load_state_dict(state_dict)


class LayerNorm(nn.Module):
    """Channel LayerNorm over (B, C, T) — parameters only; the arithmetic (modules.py:348-363 of the
    reference) is fused into tdeed_sgp_mix_fwd / tdeed_sgp_mixer_mix_fwd."""

    def __init__(self, num_channels, eps=1e-5, affine=True, device=None, dtype=None):
        super().__init__()
        factory_kwargs = {'device': device, 'dtype': dtype}
        self.num_channels = num_channels
        self.eps = eps
        self.affine = affine
        if self.affine:
            self.weight = nn.Parameter(torch.ones([1, num_channels, 1], **factory_kwargs))
            self.bias = nn.Parameter(torch.zeros([1, num_channels, 1], **factory_kwargs))
        else:
            self.register_parameter('weight', None)
            self.register_parameter('bias', None)

    def forward(self, x):
        raise NotImplementedError('LayerNorm is fused into the SGP token-mixing kernels (tdeed_sgp_mix_fwd)')


def _dw_params(mod, names):
    """{short_w, short_b} fp32 contiguous views for depthwise Conv1d modules."""
    out = {}
    for short, attr in names:
        conv = getattr(mod, attr)
        out[short + '_w'] = conv.weight.detach().float().reshape(conv.weight.shape[0], -1).contiguous()
        out[short + '_b'] = conv.bias.detach().float().contiguous()
    return out


def _mlp_forward(mod, g, y, rows, dt):
    d = y.shape[-1]
    w1 = mod.mlp[0].weight.detach().reshape(4 * d, d).to(dt).contiguous()
    w2 = mod.mlp[2].weight.detach().reshape(d, 4 * d).to(dt).contiguous()
    h = ops.gemm([(g.view(rows, d), d, 0, d)], w1, mod.mlp[0].bias.detach().float(), act=L.ACT_GELU, rows=rows)
    return ops.gemm([(h, 4 * d, 0, 4 * d)], w2, mod.mlp[2].bias.detach().float(), residual=y.reshape(rows, d), rows=rows,
                    out_dtype=torch.float32)


class SGPBlock(nn.Module):

    def __init__(self, n_embd, kernel_size=3, k=1.5, group=1, n_out=None, n_hidden=None, act_layer=nn.GELU,
                 init_conv_vars=0.1, mode='normal'):
        super().__init__()
        assert mode == 'normal', 'only the mode the reference uses is built'
        assert kernel_size % 2 == 1
        self.kernel_size = kernel_size
        if n_out is None:
            n_out = n_embd
        self.ln = LayerNorm(n_embd)
        self.gn = nn.GroupNorm(16, n_embd)
        up_size = sgp_up_size(kernel_size, k)
        self.up_size = up_size
        self.psi = nn.Conv1d(n_embd, n_embd, kernel_size, stride=1, padding=kernel_size // 2, groups=n_embd)
        self.fc = nn.Conv1d(n_embd, n_embd, 1, stride=1, padding=0, groups=n_embd)
        self.convw = nn.Conv1d(n_embd, n_embd, kernel_size, stride=1, padding=kernel_size // 2, groups=n_embd)
        self.convkw = nn.Conv1d(n_embd, n_embd, up_size, stride=1, padding=up_size // 2, groups=n_embd)
        self.global_fc = nn.Conv1d(n_embd, n_embd, 1, stride=1, padding=0, groups=n_embd)
        if n_hidden is None:
            n_hidden = 4 * n_embd
        self.mlp = nn.Sequential(
            nn.Conv1d(n_embd, n_hidden, 1, groups=group),
            act_layer(),
            nn.Conv1d(n_hidden, n_out, 1, groups=group),
        )
        self.mode = mode
        self.reset_params(init_conv_vars=init_conv_vars)

    def reset_params(self, init_conv_vars=0):
        for m in (self.psi, self.fc, self.convw, self.convkw, self.global_fc):
            torch.nn.init.normal_(m.weight, 0, init_conv_vars)
            torch.nn.init.constant_(m.bias, 0)

    def mix_weights(self):
        d = self.ln.num_channels
        w = dict(ln_w=self.ln.weight.detach().float().reshape(d).contiguous(),
                 ln_b=self.ln.bias.detach().float().reshape(d).contiguous(),
                 gn_w=self.gn.weight.detach().float().contiguous(), gn_b=self.gn.bias.detach().float().contiguous())
        w.update(_dw_params(self, (('psi', 'psi'), ('fc', 'fc'), ('convw', 'convw'), ('convkw', 'convkw'),
                                   ('gfc', 'global_fc'))))
        return w

    def forward_btc(self, x, t_out=None):
        """x: (B, T, C) fp32 CUDA; optional fused AdaptiveMaxPool1d to t_out.  -> (B, t_out, C) fp32."""
        _need_cuda(x, 'SGPBlock')
        b, t, d = x.shape
        t_out = t_out or t
        dt = _gemm_dtype()
        y, g = ops.sgp_mix(x.float().contiguous(), t_out, self.kernel_size, self.up_size, self.mix_weights(), dt)
        return _mlp_forward(self, g, y, b * t_out, dt).view(b, t_out, d)

    def forward(self, x):
        # reference layout: (B, C, T)
        return self.forward_btc(x.permute(0, 2, 1)).permute(0, 2, 1)


class SGPMixer(nn.Module):

    def __init__(self, n_embd, kernel_size=3, k=1.5, group=1, n_out=None, n_hidden=None, act_layer=nn.GELU,
                 init_conv_vars=0.1, t_size=0, concat=True):
        super().__init__()
        assert concat, 'only concat=True (the reference configuration) is built'
        assert kernel_size % 2 == 1
        self.kernel_size = kernel_size
        self.concat = concat
        self.t_size = t_size
        if n_out is None:
            n_out = n_embd
        self.ln1 = LayerNorm(n_embd)
        self.ln2 = LayerNorm(n_embd)
        self.gn = nn.GroupNorm(16, n_embd)
        up_size = sgp_up_size(kernel_size, k)
        self.up_size = up_size
        dwc = lambda ksz: nn.Conv1d(n_embd, n_embd, ksz, stride=1, padding=ksz // 2, groups=n_embd)
        self.psi1 = dwc(kernel_size)
        self.psi2 = dwc(kernel_size)
        self.convw1 = dwc(kernel_size)
        self.convkw1 = dwc(up_size)
        self.convw2 = dwc(kernel_size)
        self.convkw2 = dwc(up_size)
        self.fc1 = dwc(1)
        self.global_fc1 = dwc(1)
        self.fc2 = dwc(1)
        self.global_fc2 = dwc(1)
        if n_hidden is None:
            n_hidden = 4 * n_embd
        self.mlp = nn.Sequential(
            nn.Conv1d(n_embd, n_hidden, 1, groups=group),
            act_layer(),
            nn.Conv1d(n_hidden, n_out, 1, groups=group),
        )
        self.concat_fc = nn.Conv1d(n_embd * 6, n_embd, 1, groups=group)
        self.reset_params(init_conv_vars=init_conv_vars)

    def reset_params(self, init_conv_vars=0):
        for m in (self.psi1, self.psi2, self.convw1, self.convkw1, self.convw2, self.convkw2, self.fc1, self.fc2,
                  self.global_fc1, self.global_fc2, self.concat_fc):
            torch.nn.init.normal_(m.weight, 0, init_conv_vars)
            torch.nn.init.constant_(m.bias, 0)

    def mix_weights(self):
        d = self.ln1.num_channels
        w = dict(ln1_w=self.ln1.weight.detach().float().reshape(d).contiguous(),
                 ln1_b=self.ln1.bias.detach().float().reshape(d).contiguous(),
                 ln2_w=self.ln2.weight.detach().float().reshape(d).contiguous(),
                 ln2_b=self.ln2.bias.detach().float().reshape(d).contiguous())
        w.update(_dw_params(self, (('psi1', 'psi1'), ('psi2', 'psi2'), ('convw1', 'convw1'), ('convkw1', 'convkw1'),
                                   ('convw2', 'convw2'), ('convkw2', 'convkw2'), ('fc1', 'fc1'), ('gfc1', 'global_fc1'),
                                   ('fc2', 'fc2'), ('gfc2', 'global_fc2'))))
        return w

    def forward_btc(self, x, z):
        """x: (B, T/2, C) coarse, z: (B, T, C) skip, fp32 CUDA -> (B, T, C)."""
        _need_cuda(x, 'SGPMixer')
        b, t, d = z.shape
        dt = _gemm_dtype()
        cat = ops.sgp_mixer_mix(x.float().contiguous(), z.float().contiguous(), self.kernel_size, self.up_size,
                                self.mix_weights(), dt)
        wc = self.concat_fc.weight.detach().reshape(d, 6 * d).to(dt).contiguous()
        o = ops.gemm([(cat, 6 * d, 0, 6 * d)], wc, self.concat_fc.bias.detach().float(), act=L.ACT_GELU, rows=b * t,
                     out_dtype=torch.float32).view(b, t, d)
        g = ops.groupnorm(o, self.gn.weight.detach().float(), self.gn.bias.detach().float(), dt)
        return _mlp_forward(self, g, o, b * t, dt).view(b, t, d)

    def forward(self, x, z):
        return self.forward_btc(x.permute(0, 2, 1), z.permute(0, 2, 1)).permute(0, 2, 1)


class EDSGPMIXERLayers(nn.Module):
    def __init__(self, feat_dim, clip_len, num_layers=1, ks=3, k=2, k_factor=2, concat=True):
        super().__init__()
        self.num_layers = num_layers
        self.tot_layers = num_layers * 2 + 1
        self.clip_len = clip_len
        self.k_factor = k_factor
        self._sgp = nn.ModuleList(SGPBlock(feat_dim, kernel_size=ks, k=k, init_conv_vars=0.1)
                                  for _ in range(self.tot_layers))
        self._pooling = nn.ModuleList(nn.AdaptiveMaxPool1d(output_size=math.ceil(clip_len / (k_factor ** (i + 1))))
                                      for i in range(num_layers))
        self._sgpMixer = nn.ModuleList(SGPMixer(feat_dim, kernel_size=ks, k=k, init_conv_vars=0.1,
                                                t_size=math.ceil(clip_len / (k_factor ** i)), concat=concat)
                                       for i in range(num_layers))

    def forward(self, x):
        """x: (B, T, C) CUDA -> (B, T, C); modules.py:69-87 of the reference with pooling fused into the
        next block's kernel and the tensors kept in (B, T, C) layout throughout."""
        _need_cuda(x, 'EDSGPMIXERLayers')
        L_ = self.num_layers
        lens = [self._pooling[i - 1].output_size if i > 0 else x.shape[1] for i in range(L_ + 1)]
        x = x.float().contiguous()
        skips = []
        for i in range(L_):
            x = self._sgp[i].forward_btc(x, lens[i])
            skips.append(x)
        x = self._sgp[L_].forward_btc(x, lens[L_])
        for i in range(L_):
            x = self._sgpMixer[-(i + 1)].forward_btc(x, skips[-(i + 1)])
            x = self._sgp[L_ + i + 1].forward_btc(x)
        return x


class FCLayers(nn.Module):

    def __init__(self, feat_dim, num_classes):
        super().__init__()
        self._fc_out = nn.Linear(feat_dim, num_classes)
        self.dropout = nn.Dropout()

    def forward(self, x):
        """(B, T, C) fp32 CUDA -> logits (B, T, num_classes) through tdeed_heads_fwd (eval: dropout = identity)."""
        _need_cuda(x, 'FCLayers')
        if self.training:
            raise NotImplementedError('FCLayers standalone forward is inference-only; training goes through TDEEDModel')
        k = self._fc_out.out_features
        logits, _, _ = ops.heads(x.float().contiguous(), self._fc_out.weight.detach().float().contiguous(),
                                 self._fc_out.bias.detach().float().contiguous(), None, None, k)
        return logits


class FC2Layers(nn.Module):

    def __init__(self, feat_dim, num_classes):
        super().__init__()
        self._fc1 = FCLayers(feat_dim, num_classes[0])
        self._fc2 = FCLayers(feat_dim, num_classes[1])

    def forward(self, x):
        _need_cuda(x, 'FC2Layers')
        w = torch.cat([self._fc1._fc_out.weight, self._fc2._fc_out.weight]).detach().float().contiguous()
        b = torch.cat([self._fc1._fc_out.bias, self._fc2._fc_out.bias]).detach().float().contiguous()
        logits, _, _ = ops.heads(x.float().contiguous(), w, b, None, None, w.shape[0])
        return logits


def step(optimizer, scaler, loss, lr_scheduler=None, backward_only=False):
    if scaler is None:
        loss.backward()
    else:
        scaler.scale(loss).backward()

    if not backward_only:
        if scaler is None:
            optimizer.step()
        else:
            scaler.step(optimizer)
            scaler.update()
        if lr_scheduler is not None:
            lr_scheduler.step()
        optimizer.zero_grad()


def process_prediction(pred, predD):
    """softmax + displacement scatter-max (modules.py:406-414) as one kernel (tdeed_softmax_scatter_fwd)."""
    _need_cuda(pred, 'process_prediction')
    return ops.softmax_scatter(pred.detach().float().contiguous(), predD.detach().float().contiguous(), pred.shape[2])


def process_double_head(pred, predD, num_classes=1):
    """modules.py:416-426: softmax over the first `num_classes` columns only."""
    _need_cuda(pred, 'process_double_head')
    return ops.softmax_scatter(pred.detach().float().contiguous(), predD.detach().float().contiguous(), num_classes)


def process_labels(label, labelD, num_classes=18):
    """Host-side label bookkeeping for valMAP (modules.py:428-437); not on the device hot path."""
    label = label.cpu()
    labelD = labelD.cpu()
    label_aux = torch.zeros((label.shape[0], label.shape[1], num_classes))
    label_aux[:, :, 0] = 1
    events = label.nonzero()
    for i in range(events.shape[0]):
        b, t = int(events[i, 0]), int(events[i, 1])
        tt = t - int(labelD[b, t])
        if 0 <= tt < label.shape[1]:
            label_aux[b, tt, label[b, t]] = 1
            label_aux[b, tt, 0] = 0
    return label_aux
