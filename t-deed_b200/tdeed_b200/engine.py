"""Inference engine: T-DEED's Impl.forward(inference=True) (model/model.py:105-149 of the reference)
as a fixed sequence of libtdeed_sm100 kernel launches.

Host responsibilities (Python, as the north star prescribes): fold eval-mode BatchNorm into the
conv weights, lay weights out for the kernels (1x1 convs as K-major [N, K] matrices, bf16 for the
tcgen05 path), own device buffers, issue launches on the current stream, and optionally capture
the whole forward in a CUDA graph (static shapes: B clips x T frames x H x W).

precision: 'bf16' = bf16 activations/weights in the backbone and for every GEMM operand, fp32
accumulation, fp32 SGP residual stream (the reference runs fp16 autocast); 'fp32' = exact fp32
CUDA-core path used for 1e-3 parity against the reference's fp32 forward.
"""
import math

import torch

from . import _lib as L
from . import ops

REGNET = {
    'rny002': dict(widths=[24, 56, 152, 368], depths=[1, 1, 4, 7], group_width=8),
    'rny008': dict(widths=[64, 128, 320, 768], depths=[1, 3, 8, 2], group_width=16),
}


def fold_dim(channels, n_div=4):
    """model/shift.py:79 of the reference."""
    return math.ceil(channels // n_div / 4) * 4


def sgp_up_size(ks, k):
    """model/modules.py:119-120 of the reference."""
    up = round((ks + 1) * k)
    return up + 1 if up % 2 == 0 else up


def center_crop_offsets(h, w, crop):
    """torchvision CenterCrop offsets (model/model.py:100,124)."""
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def _round8(v):
    return (v + 7) // 8 * 8


class EngineConfig:
    def __init__(self, feature_arch, clip_len, n_layers, sgp_ks, sgp_r, num_classes, radi_displacement, crop_dim,
                 double_head=None):
        self.backbone, self.shift_mode = feature_arch.rsplit('_', 1) if '_' in feature_arch else (feature_arch, None)
        if self.backbone not in REGNET:
            raise NotImplementedError(feature_arch)
        self.clip_len = clip_len
        self.n_layers = n_layers
        self.sgp_ks = sgp_ks
        self.sgp_up = sgp_up_size(sgp_ks, sgp_r)
        self.num_classes = num_classes
        self.radi_displacement = radi_displacement
        self.crop_dim = crop_dim if (crop_dim is not None and crop_dim > 0) else None
        self.double_head = list(double_head) if double_head else None
        self.feat_dim = REGNET[self.backbone]['widths'][-1]

    def blocks(self):
        r = REGNET[self.backbone]
        prev = 32
        for si, (w, d) in enumerate(zip(r['widths'], r['depths'])):
            for bi in range(d):
                yield ('_features.s%d.b%d' % (si + 1, bi + 1), prev if bi == 0 else w, w, 2 if bi == 0 else 1,
                       si >= 2 and self.shift_mode in ('gsm', 'gsf'))
            prev = w


PAD_LIMIT = 1.07      # 56 -> 64 (stage 2, +14 % bytes) was measured too: no gain


def padded_width(w):
    """bf16 rows of `w` channels are w*2 bytes; the GEMM / grouped-conv epilogues store 32 bytes per instruction when every row
    starts on a 32-byte boundary (w % 16 == 0).  Widths that miss it by <= 7 % are zero-padded (RegNetY-200MF stage 3: 152 -> 160,
    +5 % bytes, conv1 GEMM 218 -> 145 us per 57-clip batch); the pad channels carry exact zeros through the whole block."""
    wp = (w + 15) // 16 * 16
    return wp if wp != w and wp <= PAD_LIMIT * w else w


def _pad_to(t, shape, fill=0.0):
    out = torch.full(shape, fill, dtype=t.dtype)
    out[tuple(slice(0, n) for n in t.shape)] = t
    return out


def pad_block_state(sd, p, cin, cout, cin_p, cout_p, shifted):
    """Zero-pad the tensors of bottleneck `p` from (cin, cout) to (cin_p, cout_p) channels: padded output channels get zero
    weights, zero BN scale / shift (weight 0, bias 0, mean 0, var 1) and zero SE columns, padded input channels zero columns."""
    if (cin, cout) == (cin_p, cout_p):
        return

    def bn(q, n):
        for k, fill in (('weight', 0.0), ('bias', 0.0), ('running_mean', 0.0), ('running_var', 1.0)):
            sd[q + '.' + k] = _pad_to(sd[q + '.' + k], (n,), fill)

    c1 = p + ('.conv1.net' if shifted else '.conv1')
    sd[c1 + '.conv.weight'] = _pad_to(sd[c1 + '.conv.weight'], (cout_p, cin_p, 1, 1))
    bn(c1 + '.bn', cout_p)
    w2 = sd[p + '.conv2.conv.weight']
    sd[p + '.conv2.conv.weight'] = _pad_to(w2, (cout_p,) + tuple(w2.shape[1:]))
    bn(p + '.conv2.bn', cout_p)
    rd = sd[p + '.se.fc1.weight'].shape[0]
    sd[p + '.se.fc1.weight'] = _pad_to(sd[p + '.se.fc1.weight'], (rd, cout_p, 1, 1))
    sd[p + '.se.fc2.weight'] = _pad_to(sd[p + '.se.fc2.weight'], (cout_p, rd, 1, 1))
    sd[p + '.se.fc2.bias'] = _pad_to(sd[p + '.se.fc2.bias'], (cout_p,))
    sd[p + '.conv3.conv.weight'] = _pad_to(sd[p + '.conv3.conv.weight'], (cout_p, cout_p, 1, 1))
    bn(p + '.conv3.bn', cout_p)
    if (p + '.downsample.conv.weight') in sd:
        sd[p + '.downsample.conv.weight'] = _pad_to(sd[p + '.downsample.conv.weight'], (cout_p, cin_p, 1, 1))
        bn(p + '.downsample.bn', cout_p)


def _bn_fold(sd, p, eps=1e-5):
    scale = sd[p + '.weight'].float() / torch.sqrt(sd[p + '.running_var'].float() + eps)
    shift = sd[p + '.bias'].float() - sd[p + '.running_mean'].float() * scale
    return scale, shift


class InferenceEngine:
    """Prepared weights + launch sequence.  `state` is a state_dict with the reference's key names."""

    def __init__(self, cfg, state, device='cuda', precision='bf16', gemm_backend=L.GEMM_AUTO):
        if not torch.cuda.is_available():
            raise RuntimeError('tdeed_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        L.load()
        assert precision in ('bf16', 'fp32')
        self.cfg = cfg
        self.device = torch.device(device)
        self.precision = precision
        self.act_dtype = torch.bfloat16 if precision == 'bf16' else torch.float32
        self.gemm_backend = gemm_backend
        self.launches = 0
        self.prof = None
        self.fuse_stem = True            # bf16: tcgen05 stem fused with s1.b1.conv1
        self.conv3_tc = True             # bf16: grouped 3x3 conv on tcgen05
        self.fuse_ds0 = True             # bf16: s1.b1 shortcut conv as a second K-segment of conv3's GEMM
        self.stem_v2 = True              # bf16 + uint8 frames: raw-pixel shifted-descriptor stem (stem_tc2.cu)
        self.fuse_se = True              # bf16: SE gate folded into conv3's A operand (stages 3-4)
        self._graphs = {}
        self.load_state(state)

    # ------------------------------------------------------------------ weights
    def load_state(self, state):
        # Weight preparation (BatchNorm folding, K-segment padding, UMMA weight images) is host arithmetic on CPU tensors, uploaded
        # once at the end: no eager-torch GPU kernels (r1 launched ~600 ATen elementwise kernels per engine build, which
        # drowned the library's own kernels in the driver's launch list).
        cfg, adt = self.cfg, self.act_dtype
        dev = torch.device('cpu')
        sd = {k: v.detach().cpu() for k, v in state.items()}
        f32 = lambda t: t.float().contiguous()
        W = {}
        sc, sh = _bn_fold(sd, '_features.stem.bn')
        W['stem_w'] = f32(sd['_features.stem.conv.weight'].float() * sc[:, None, None, None])
        W['stem_b'] = f32(sh)
        w0 = torch.zeros((32, 32), dtype=torch.float32, device=dev)
        w0[:, :27] = W['stem_w'].reshape(32, 27)
        W['stem_w_tc'] = w0.to(torch.bfloat16).contiguous()        # tcgen05 stem: K = 27 padded to 32
        W['stem2_wimg'], W['stem2_b0'], W['stem2_pad'] = ops.stem_tc2_weights(W['stem_w'], W['stem_b'])   # raw-pixel stem (stem_tc2.cu)
        gw = REGNET[cfg.backbone]['group_width']
        blocks = []
        feat = cfg.feat_dim
        for p, cin, cout, stride, shifted in cfg.blocks():
            fd_real = fold_dim(cin)
            if adt == torch.bfloat16:
                # 32-byte aligned activation rows for the inner stages (the last stage's width is the feature dimension)
                cin_p = padded_width(cin) if cin != 32 else cin
                cout_p = padded_width(cout) if cout != feat else cout
                assert fold_dim(cin_p) == fd_real or not shifted
                pad_block_state(sd, p, cin, cout, cin_p, cout_p, shifted)
                cin, cout = cin_p, cout_p
            b = dict(cin=cin, cout=cout, stride=stride, shifted=shifted, gw=gw)
            c1 = p + ('.conv1.net' if shifted else '.conv1')
            sc, sh = _bn_fold(sd, c1 + '.bn')
            w1 = sd[c1 + '.conv.weight'].float().reshape(cout, cin) * sc[:, None]
            if shifted:
                # K layout of the virtual concat [gate-shift out (fold, padded to 8) | x[:, xs:]]: TMA box origins
                # must be 16-byte aligned, so the x segment starts at the multiple of 8 below `fold`; the few
                # overlapped channels / pad columns get zero weights.
                fd = fd_real
                fdp, xs = _round8(fd), fd // 8 * 8
                wp = torch.zeros((cout, fdp + cin - xs), dtype=torch.float32, device=dev)
                # the gate-shift kernel writes its channels in input order (ops.gsf natural=True): the reference's channel
                # interleave (model/impl/gsf.py:84-92) becomes a column permutation of this weight
                wp[:, :fd] = w1[:, ops.gsf_interleaved_positions(fd)]
                wp[:, fdp + (fd - xs):] = w1[:, fd:]
                w1 = wp
                b['x_start'] = xs
            b['w1'] = w1.to(adt).contiguous()
            b['b1'] = f32(sh)
            sc, sh = _bn_fold(sd, p + '.conv2.bn')
            b['w2'] = f32(sd[p + '.conv2.conv.weight'].float() * sc[:, None, None, None])
            b['b2'] = f32(sh)
            if adt == torch.bfloat16:
                b['w2_img'] = ops.conv3_weight_image(b['w2'], gw)      # tcgen05 B tiles
            rd = sd[p + '.se.fc1.weight'].shape[0]
            b['se_w1'] = f32(sd[p + '.se.fc1.weight'].reshape(rd, cout))
            b['se_b1'] = f32(sd[p + '.se.fc1.bias'])
            b['se_w2'] = f32(sd[p + '.se.fc2.weight'].reshape(cout, rd).t())     # transposed: [rd][c]
            b['se_b2'] = f32(sd[p + '.se.fc2.bias'])
            sc, sh = _bn_fold(sd, p + '.conv3.bn')
            b['w3'] = (sd[p + '.conv3.conv.weight'].float().reshape(cout, cout) * sc[:, None]).to(adt).contiguous()
            b['b3'] = f32(sh)
            if (p + '.downsample.conv.weight') in sd:
                sc, sh = _bn_fold(sd, p + '.downsample.bn')
                b['wd'] = (sd[p + '.downsample.conv.weight'].float().reshape(cout, cin) * sc[:, None]).to(adt).contiguous()
                b['bd'] = f32(sh)
            if shifted:
                g = p + '.conv1.gs'
                fd = fd_real
                sc, sh = _bn_fold(sd, g + '.bn')
                gs = dict(fold=fd, bn_scale=f32(sc), bn_shift=f32(sh), w3d=f32(sd[g + '.conv3D.weight'].reshape(-1)),
                          b3d=f32(sd[g + '.conv3D.bias']))
                if cfg.shift_mode == 'gsf':
                    gs['cc_w'] = f32(torch.cat([sd[g + '.channel_conv1.weight'].reshape(-1),
                                                sd[g + '.channel_conv2.weight'].reshape(-1)]))
                    gs['cc_b'] = f32(torch.cat([sd[g + '.channel_conv1.bias'], sd[g + '.channel_conv2.bias']]))
                b['gs'] = gs
            if not blocks and not shifted and cin == 32 and cout <= 64:
                # s1.b1.conv1 fused into the tensor-core stem: rows padded to a multiple of 16
                wf = torch.zeros(((cout + 15) // 16 * 16, 32), dtype=torch.float32, device=dev)
                wf[:cout] = w1
                b['w1_fused'] = wf.to(torch.bfloat16).contiguous()
                if 'wd' in b and adt == torch.bfloat16:
                    # the stem kernel writes the stride-2 subsample of its output compactly, so the shortcut conv of s1.b1 is a
                    # second K-segment of conv3's GEMM:  relu([a2 | x_sub] @ [W3 | Wd]^T + b3 + bd)  — one launch, no `res` tensor
                    b['w3d'] = torch.cat([b['w3'].float(), b['wd'].float()], dim=1).to(adt).contiguous()
                    b['b3d'] = (b['b3'] + b['bd']).contiguous()
            blocks.append(b)
        W['blocks'] = blocks
        W['temp_enc'] = f32(sd['temp_enc'])
        d = cfg.feat_dim

        def dw(name):
            return f32(sd[name + '.weight'].reshape(d, -1)), f32(sd[name + '.bias'])

        def mlp(p):
            return dict(w1=sd[p + '.mlp.0.weight'].reshape(4 * d, d).to(adt).contiguous(), b1=f32(sd[p + '.mlp.0.bias']),
                        w2=sd[p + '.mlp.2.weight'].reshape(d, 4 * d).to(adt).contiguous(), b2=f32(sd[p + '.mlp.2.bias']))

        sgp = []
        for i in range(2 * cfg.n_layers + 1):
            p = '_temp_fine._sgp.%d' % i
            w = dict(ln_w=f32(sd[p + '.ln.weight'].reshape(d)), ln_b=f32(sd[p + '.ln.bias'].reshape(d)),
                     gn_w=f32(sd[p + '.gn.weight']), gn_b=f32(sd[p + '.gn.bias']))
            for n, key in (('psi', 'psi'), ('fc', 'fc'), ('convw', 'convw'), ('convkw', 'convkw'), ('gfc', 'global_fc')):
                w[n + '_w'], w[n + '_b'] = dw(p + '.' + key)
            sgp.append(dict(mix=w, mlp=mlp(p)))
        W['sgp'] = sgp
        mixers = []
        for i in range(cfg.n_layers):
            p = '_temp_fine._sgpMixer.%d' % i
            w = dict(ln1_w=f32(sd[p + '.ln1.weight'].reshape(d)), ln1_b=f32(sd[p + '.ln1.bias'].reshape(d)),
                     ln2_w=f32(sd[p + '.ln2.weight'].reshape(d)), ln2_b=f32(sd[p + '.ln2.bias'].reshape(d)))
            for n, key in (('psi1', 'psi1'), ('psi2', 'psi2'), ('convw1', 'convw1'), ('convkw1', 'convkw1'),
                           ('convw2', 'convw2'), ('convkw2', 'convkw2'), ('fc1', 'fc1'), ('gfc1', 'global_fc1'),
                           ('fc2', 'fc2'), ('gfc2', 'global_fc2')):
                w[n + '_w'], w[n + '_b'] = dw(p + '.' + key)
            mixers.append(dict(mix=w, mlp=mlp(p), gn_w=f32(sd[p + '.gn.weight']), gn_b=f32(sd[p + '.gn.bias']),
                               wc=sd[p + '.concat_fc.weight'].reshape(d, 6 * d).to(adt).contiguous(),
                               bc=f32(sd[p + '.concat_fc.bias'])))
        W['mixers'] = mixers
        if cfg.double_head:
            W['cls_w'] = f32(torch.cat([sd['_pred_fine._fc1._fc_out.weight'], sd['_pred_fine._fc2._fc_out.weight']]))
            W['cls_b'] = f32(torch.cat([sd['_pred_fine._fc1._fc_out.bias'], sd['_pred_fine._fc2._fc_out.bias']]))
        else:
            W['cls_w'] = f32(sd['_pred_fine._fc_out.weight'])
            W['cls_b'] = f32(sd['_pred_fine._fc_out.bias'])
        if cfg.radi_displacement > 0:
            W['displ_w'] = f32(sd['_pred_displ._fc_out.weight'].reshape(-1))
            W['displ_b'] = f32(sd['_pred_displ._fc_out.bias'])
        else:
            W['displ_w'] = W['displ_b'] = None
        def upload(o):
            if isinstance(o, torch.Tensor):
                return o.to(self.device)
            if isinstance(o, dict):
                return {k: upload(v) for k, v in o.items()}
            if isinstance(o, list) and o and isinstance(o[0], (dict, torch.Tensor)):
                return [upload(v) for v in o]
            return o
        self.W = upload(W)
        self._graphs.clear()
        self.version = getattr(self, 'version', 0) + 1      # bumps whenever prepared weights change (caches keyed on it)

    # ------------------------------------------------------------------ forward
    def _op(self, label, flops, nbytes, fn, *a, **k):
        """Launch one kernel family; when self.prof is a list, bracket it with CUDA events and record the
        algorithmic FLOPs / bytes for bench.py."""
        self.launches += k.pop('_n', 1)
        if self.prof is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        self.prof.append((label, flops, nbytes, e0, e1))
        return out

    def _gemm(self, segs, w, bias=None, label='gemm', **k):
        rows = k['rows']
        n, kk = w.shape
        es = w.element_size()
        out_es = 4 if k.get('out_dtype') == torch.float32 else es
        nbytes = rows * kk * es + n * kk * es + rows * n * out_es + (rows * n * k['residual'].element_size() if k.get('residual') is not None else 0)
        return self._op(label, 2.0 * rows * n * kk, nbytes, ops.gemm, segs, w, bias, backend=self.gemm_backend, **k)

    def crop_window(self, in_h, in_w):
        if self.cfg.crop_dim is None:
            return 0, 0, in_h, in_w
        cy, cx = center_crop_offsets(in_h, in_w, self.cfg.crop_dim)
        return cy, cx, self.cfg.crop_dim, self.cfg.crop_dim

    def split_index(self):
        """Number of leading bottleneck blocks that are clip-independent: everything before the first GatedShift
        (model/shift.py:47-59 wraps conv1 of every block of s3 and s4).  Their per-frame output can be cached and
        shared by the overlapping clips of a video (see tdeed_b200.pipeline.VideoInference)."""
        for i, blk in enumerate(self.W['blocks']):
            if blk['shifted']:
                return i
        return len(self.W['blocks'])

    def lower(self, frames, flip=False, crop=None, taps=None):
        """frames (N,3,H,W) u8|f32 on device -> NHWC activation (N,h,w,c) after the last clip-independent block
        (stem, s1, s2 for the gate-shift backbones).  No cross-frame arithmetic: frame i of the output depends on frame i
        of the input only, so the result is the same whichever clip / batch a frame arrives in."""
        W, adt = self.W, self.act_dtype
        n = frames.shape[0]
        in_h, in_w = frames.shape[-2:]
        crop = crop or self.crop_window(in_h, in_w)
        es = 2 if adt == torch.bfloat16 else 4
        oh0, ow0 = (crop[2] + 1) // 2, (crop[3] + 1) // 2
        fuse = adt == torch.bfloat16 and self.fuse_stem and 'w1_fused' in W['blocks'][0] and taps is None
        a1_fused = None
        if fuse:
            # tcgen05 stem + s1.b1.conv1 in one kernel; only the stride-2 subsample of the stem output (the
            # shortcut conv's input) and conv1's output reach HBM
            c1 = W['blocks'][0]['cout']
            if frames.dtype == torch.uint8 and self.stem_v2:
                x_sub, a1_fused = self._op('stem', 2.0 * n * oh0 * ow0 * 32 * (27 + c1),
                                           n * (3 * crop[2] * crop[3] * frames.element_size() + oh0 * ow0 * (c1 + 8) * es),
                                           ops.stem_tc2, frames, crop, flip, W['stem2_wimg'], W['stem2_b0'],
                                           W['stem2_pad'], W['blocks'][0]['w1_fused'], W['blocks'][0]['b1'], c1, 2)
            else:
                x_sub, a1_fused = self._op('stem', 2.0 * n * oh0 * ow0 * 32 * (27 + c1),
                                           n * (3 * crop[2] * crop[3] * frames.element_size() + oh0 * ow0 * (c1 + 8) * es),
                                           ops.stem_tc, frames, crop, flip, W['stem_w_tc'], W['stem_b'],
                                           W['blocks'][0]['w1_fused'], W['blocks'][0]['b1'], c1, True, 2)
            x = None
        elif adt == torch.bfloat16 and self.fuse_stem:
            x, _ = self._op('stem', 2.0 * n * oh0 * ow0 * 32 * 27, n * (3 * crop[2] * crop[3] * frames.element_size() + oh0 * ow0 * 32 * es),
                            ops.stem_tc, frames, crop, flip, W['stem_w_tc'], W['stem_b'])
        else:
            x = self._op('stem', 2.0 * n * oh0 * ow0 * 32 * 27, n * (3 * crop[2] * crop[3] * frames.element_size() + oh0 * ow0 * 32 * es),
                         ops.stem, frames, crop, flip, W['stem_w'], W['stem_b'], adt)
        if taps is not None:
            taps['stem'] = x
        for bi in range(self.split_index()):
            blk = W['blocks'][bi]
            if bi == 0 and a1_fused is not None:
                x = self._block(bi, blk, None, n, None, None, taps, fused=(a1_fused, x_sub, oh0, ow0))
            else:
                x = self._block(bi, blk, x, n, None, None, taps)
        return x

    def upper(self, x, b, t, taps=None):
        """x NHWC (b*t, h, w, c): output of lower() laid out clip-major -> feat (b*t, d) fp32 (pooled + temp_enc)."""
        W = self.W
        n = b * t
        es = 2 if self.act_dtype == torch.bfloat16 else 4
        for bi in range(self.split_index(), len(W['blocks'])):
            x = self._block(bi, W['blocks'][bi], x, n, b, t, taps)
        hw = x.shape[1] * x.shape[2]
        return self._op('pool_posenc', float(n * hw * x.shape[3]), n * hw * x.shape[3] * es + n * x.shape[3] * 4,
                        ops.pool_posenc, x, t, W['temp_enc'])

    def _block(self, bi, blk, x, n, b, t, taps, fused=None):
        """One RegNetY bottleneck incl. its gate-shift prologue (a5-a7) when `shifted`."""
        cfg, adt = self.cfg, self.act_dtype
        es = 2 if adt == torch.bfloat16 else 4
        if fused is not None:
            a1_fused, x_sub, h, w = fused
            cin = 32
        else:
            a1_fused = None
            _, h, w, cin = x.shape
        cout, stride = blk['cout'], blk['stride']
        m = n * h * w
        if a1_fused is not None:
            a1 = a1_fused
        else:
            if blk['shifted']:
                gs = blk['gs']
                fd = gs['fold']
                ws = torch.empty(ops.gsf_workspace_floats(b, t, h, w, fd), dtype=torch.float32, device=x.device)
                gso = torch.empty((m, _round8(fd)), dtype=adt, device=x.device)
                self._op('gsf', 2.0 * m * 27 * fd, m * fd * es * 2, ops.gsf, x, b, t, fd,
                         L.SHIFT_GSF if cfg.shift_mode == 'gsf' else L.SHIFT_GSM, gs, ws, gso, natural=True,
                         _n=4 if cfg.shift_mode == 'gsf' else 3)
                segs = [(gso, gso.shape[1], 0, gso.shape[1]), (x, cin, blk['x_start'], cin - blk['x_start'])]
            else:
                segs = [(x, cin, 0, cin)]
            a1 = self._gemm(segs, blk['w1'], blk['b1'], label='conv1x1', act=L.ACT_RELU, rows=m).view(n, h, w, cout)
        oh, ow = (h + stride - 1) // stride, (w + stride - 1) // stride
        mo = n * oh * ow
        if 'w2_img' in blk and self.conv3_tc:
            a2 = self._op('conv3x3g', 2.0 * mo * cout * 9 * blk['gw'], (m + mo) * cout * es + cout * 9 * blk['gw'] * 4,
                          ops.conv3x3g_tc, a1, blk['w2_img'], blk['b2'], stride)
        else:
            a2 = self._op('conv3x3g', 2.0 * mo * cout * 9 * blk['gw'], (m + mo) * cout * es + cout * 9 * blk['gw'] * 4,
                          ops.conv3x3g, a1, blk['w2'], blk['b2'], blk['gw'], stride)
        # squeeze-excite.  K > 64 layers on the tcgen05 path (stages 3-4): only the gate is computed here (mean + fc) and conv3's
        # GEMM applies it to its A operand in shared memory — no read-modify-write pass over a2.  Elsewhere: in place.
        # (measured, 57-clip batch: N = K = 152 conv3 222 + 117 us (se_scale) -> 260 us fused; N = K = 368: 176 + 72 -> 281 us, a
        # loss — two n-tiles scale every A block twice behind a 3-stage ring — so the fold is taken for one-n-tile layers only)
        fold_se = (self.fuse_se and adt == torch.bfloat16 and self.gemm_backend in (L.GEMM_AUTO, L.GEMM_TCGEN05) and 64 < cout <= 256
                   and cout % 8 == 0 and not (a1_fused is not None and 'w3d' in blk and self.fuse_ds0))
        gate = None
        if fold_se:
            gate = self._op('se', 4.0 * n * cout * blk['se_w1'].shape[0], mo * cout * es,
                            ops.se_gate, a2, blk['se_w1'], blk['se_b1'], blk['se_w2'], blk['se_b2'], _n=2)
        else:
            self._op('se', 4.0 * n * cout * blk['se_w1'].shape[0], 2 * mo * cout * es,
                     ops.se_, a2, blk['se_w1'], blk['se_b1'], blk['se_w2'], blk['se_b2'], _n=3)
        if a1_fused is not None and 'w3d' in blk and self.fuse_ds0:
            x = self._gemm([(a2.view(mo, cout), cout, 0, cout), (x_sub.view(mo, 32), 32, 0, 32)], blk['w3d'], blk['b3d'],
                           label='conv1x1', act=L.ACT_RELU, rows=mo).view(n, oh, ow, cout)
        else:
            if a1_fused is not None:
                # the stem kernel already wrote the stride-2 subsample: plain GEMM, no gather
                res = self._gemm([(x_sub, 32, 0, 32)], blk['wd'], blk['bd'], label='conv1x1_ds', rows=mo)
            elif 'wd' in blk:
                res = self._gemm([(x, cin, 0, cin)], blk['wd'], blk['bd'], label='conv1x1_ds', rows=mo,
                                 gather=(stride, h, w) if stride > 1 else None)
            else:
                res = x.view(mo, cout)
            if gate is not None:
                es3 = blk['w3'].element_size()
                x = self._op('conv1x1', 2.0 * mo * cout * cout, (mo * cout * 3 + cout * cout) * es3,
                             ops.gemm_scaled, a2.view(mo, cout), gate, oh * ow, blk['w3'], blk['b3'], residual=res,
                             act=L.ACT_RELU).view(n, oh, ow, cout)
            else:
                x = self._gemm([(a2.view(mo, cout), cout, 0, cout)], blk['w3'], blk['b3'], label='conv1x1', residual=res,
                               act=L.ACT_RELU, rows=mo).view(n, oh, ow, cout)
        if taps is not None:
            taps['s%d.b%d' % (self._stage_of(bi))] = x
        return x

    def backbone(self, frames, flip=False, crop=None, taps=None):
        """frames (B,T,3,H,W) u8|f32 on device -> feat (B*T, d) fp32 (pooled + temp_enc)."""
        b, t = frames.shape[:2]
        x = self.lower(frames.reshape(b * t, 3, frames.shape[-2], frames.shape[-1]), flip=flip, crop=crop, taps=taps)
        return self.upper(x, b, t, taps=taps)

    def _stage_of(self, bi):
        depths = REGNET[self.cfg.backbone]['depths']
        s = 0
        while bi >= depths[s]:
            bi -= depths[s]
            s += 1
        return s + 1, bi + 1

    def _mlp(self, g, y, p, rows):
        h = self._gemm([(g.view(rows, -1), g.shape[-1], 0, g.shape[-1])], p['w1'], p['b1'], label='sgp_gemm',
                       act=L.ACT_GELU, rows=rows)
        return self._gemm([(h, h.shape[1], 0, h.shape[1])], p['w2'], p['b2'], label='sgp_gemm', residual=y.view(rows, -1),
                          rows=rows, out_dtype=torch.float32)

    def _sgp_block(self, x, t_out, p):
        cfg = self.cfg
        b, _, d = x.shape
        y, g = self._op('sgp_mix', 2.0 * b * t_out * d * (2 * cfg.sgp_ks + cfg.sgp_up + 2), b * d * (x.shape[1] * 4 + t_out * 6),
                        ops.sgp_mix, x, t_out, cfg.sgp_ks, cfg.sgp_up, p['mix'], self.act_dtype)
        return self._mlp(g, y, p['mlp'], b * t_out).view(b, t_out, d)

    def temporal(self, feat):
        """(B,T,d) fp32 -> (B,T,d) fp32: EDSGPMIXERLayers.forward (model/modules.py:69-87)."""
        cfg, W = self.cfg, self.W
        b, t, d = feat.shape
        L_ = cfg.n_layers
        lens = [math.ceil(t / 2 ** i) for i in range(L_ + 1)]
        x = feat
        skips = []
        for i in range(L_):
            x = self._sgp_block(x, lens[i], W['sgp'][i])     # pooling of the previous level is fused here
            skips.append(x)
        x = self._sgp_block(x, lens[L_], W['sgp'][L_])
        for i in range(L_):
            j = L_ - 1 - i
            mp = W['mixers'][j]
            rows = b * lens[j]
            cat = self._op('sgp_mix', 4.0 * rows * d * (2 * cfg.sgp_ks + cfg.sgp_up + 2), rows * d * (6 + 12),
                           ops.sgp_mixer_mix, x, skips[j], cfg.sgp_ks, cfg.sgp_up, mp['mix'], self.act_dtype)
            o = self._gemm([(cat, 6 * d, 0, 6 * d)], mp['wc'], mp['bc'], label='sgp_gemm', act=L.ACT_GELU, rows=rows,
                           out_dtype=torch.float32).view(b, lens[j], d)
            g = self._op('sgp_mix', 8.0 * rows * d, rows * d * 6, ops.groupnorm, o, mp['gn_w'], mp['gn_b'], self.act_dtype)
            x = self._mlp(g, o, mp['mlp'], rows).view(b, lens[j], d)
            x = self._sgp_block(x, lens[j], W['sgp'][L_ + i + 1])
        return x

    def heads(self, x):
        cfg, W = self.cfg, self.W
        b, t, d = x.shape
        return self._op('heads', 2.0 * b * t * d * (W['cls_w'].shape[0] + 1), b * t * d * 4,
                        ops.heads, x, W['cls_w'], W['cls_b'], W['displ_w'], W['displ_b'], cfg.num_classes + 1)

    def forward(self, frames, flip=False, crop=None, taps=None):
        """frames (B,T,3,H,W) u8|f32 device tensor -> (logits (B,T,K'), displ (B,T)|None, probs (B,T,K))."""
        b, t = frames.shape[:2]
        feat = self.backbone(frames, flip=flip, crop=crop, taps=taps).view(b, t, self.cfg.feat_dim)
        if taps is not None:
            taps['feat_posenc'] = feat
        x = self.temporal(feat)
        if taps is not None:
            taps['temporal'] = x
        return self.heads(x)

    # ------------------------------------------------------------------ CUDA graph replay
    def graphed(self, key, fn):
        """Run `fn()` — a closure that launches engine kernels reading STATIC device tensors — through a CUDA graph captured
        the first time `key` is seen (two eager warm-up runs on a side stream set kernel attributes and prime the allocator).
        Returns fn's outputs: views of graph-owned buffers, overwritten by the next replay of the same key."""
        ent = self._graphs.get(key)
        if ent is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = self.launches
            # thread_local: a DataLoader pin-memory thread may be calling into CUDA while we capture
            with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                outs = fn()
            ent = dict(graph=graph, outs=outs, launches=self.launches - n0)
            self._graphs[key] = ent
        ent['graph'].replay()
        self.launches += ent['launches']
        return ent['outs']

    def forward_graphed(self, frames, flip=False, crop=None, static=False):
        """forward() replayed from a CUDA graph (static shapes).  static=False: `frames` is copied into a graph-owned input
        buffer first (any tensor works).  static=True: the graph is captured on `frames`' own storage — the caller promises
        that the same buffer (e.g. one of ClipUploader's two) is refilled in place for later calls; no extra copy."""
        shape_key = (tuple(frames.shape), frames.dtype, bool(flip), crop)
        if static:
            return self.graphed(('fwd', frames.data_ptr()) + shape_key, lambda: self.forward(frames, flip=flip, crop=crop))
        buf = self._graphs.get(('fwd_in',) + shape_key)
        if buf is None:
            buf = self._graphs[('fwd_in',) + shape_key] = torch.empty_like(frames)
        buf.copy_(frames, non_blocking=True)
        return self.graphed(('fwd', buf.data_ptr()) + shape_key, lambda: self.forward(buf, flip=flip, crop=crop))

    def lower_graphed(self, frames, flip=False, crop=None):
        """lower() on a static (N,3,H,W) device buffer, replayed from a CUDA graph keyed by the buffer's address."""
        key = ('lower', frames.data_ptr(), tuple(frames.shape), frames.dtype, bool(flip), crop)
        return self.graphed(key, lambda: self.lower(frames, flip=flip, crop=crop))

    def upper_feat_graphed(self, x, b, t):
        """upper() alone on a static clip-major feature buffer: the pooled per-frame features (b*t, feat_dim) fp32."""
        key = ('upper_feat', x.data_ptr(), tuple(x.shape), b, t)
        return self.graphed(key, lambda: self.upper(x, b, t))

    def temporal_heads_graphed(self, feat, b, t):
        """temporal() + heads() on a static (b*t, feat_dim) feature buffer — the clips of several backbone batches at once: the
        temporal stack is a few dozen small launches whose time barely depends on b."""
        key = ('temporal', feat.data_ptr(), b, t)
        return self.graphed(key, lambda: self.heads(self.temporal(feat.view(b, t, self.cfg.feat_dim))))

    def upper_graphed(self, x, b, t):
        """upper() + temporal() + heads() on a static clip-major feature buffer x (b*t, h, w, c)."""
        key = ('upper', x.data_ptr(), tuple(x.shape), b, t)
        return self.graphed(key, lambda: self.heads(self.temporal(self.upper(x, b, t).view(b, t, self.cfg.feat_dim))))
