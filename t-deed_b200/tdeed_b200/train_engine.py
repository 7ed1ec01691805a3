"""Training step on the C-ABI kernels: forward in train mode (batch-statistics BatchNorm, dropout), loss, and a
hand-written backward pass that writes every parameter gradient — no torch.autograd, no cuDNN, no CPU fallback.

Reference: TDEEDModel.epoch with an optimizer (model/model.py:193-332) = Impl.forward(inference=False) (:105-149) +
F.cross_entropy / F.mse_loss (:308-319) + loss.backward() (model/modules.py:388-401).

Layout: activations NHWC in `act_dtype` (bf16 by default, fp32 for the parity mode); the temporal layers and all
parameter gradients are fp32.  Parameters are read from `params` (name -> fp32 device tensor, the reference's
state-dict names); gradients are written into `grads` (same names / shapes).  GEMM operands in bf16 come from
`shadow` (name -> bf16 copy maintained by the fused AdamW kernel) when available.
"""
import math

import torch

from . import _lib as L
from . import ops
from . import train_ops as T
from .engine import REGNET, fold_dim, sgp_up_size


class TrainEngine:
    def __init__(self, cfg, params, buffers, grads, act_dtype=torch.bfloat16, shadow=None, gemm_backend=L.GEMM_AUTO):
        if not torch.cuda.is_available():
            raise RuntimeError('tdeed_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        L.load()
        if cfg.shift_mode not in ('gsf', 'gsm'):
            raise NotImplementedError('training needs a gate-shift backbone (rny00X_gsf / _gsm)')
        self.cfg = cfg
        self.P, self.B, self.G = params, buffers, grads
        self.adt = act_dtype
        self.shadow = shadow
        self.gemm_backend = gemm_backend if act_dtype == torch.bfloat16 else L.GEMM_SIMT
        self.gw = REGNET[cfg.backbone]['group_width']
        self.launch_log = None
        self._class_weights = {}
        dev = next(iter(params.values())).device
        self.seed_dev = torch.full((1,), torch.initial_seed() % (1 << 30), dtype=torch.int64, device=dev)

    # ------------------------------------------------------------------ helpers
    def _wop(self, name, rows, cols):
        """[rows, cols] GEMM operand of a weight in the activation dtype (+ its transpose for the data gradient)."""
        if self.adt == torch.float32:
            w = self.P[name].reshape(rows, cols)
        elif self.shadow is not None:
            w = self.shadow[name].reshape(rows, cols)
        else:
            w = T.cast(self.P[name].reshape(rows, cols), torch.bfloat16)
        return w, w.t().contiguous()

    def _gemm(self, a2d, w, rows, bias=None, residual=None, out_dtype=None, gather=None):
        return ops.gemm([(a2d, a2d.stride(0), 0, w.shape[1])], w, bias, residual=residual, rows=rows, out_dtype=out_dtype,
                        gather=gather, backend=self.gemm_backend)

    def _bn(self, y2d, C, p, M=None):
        self.B[p + '.num_batches_tracked'].add_(1)
        return T.bn_stats(y2d, C, self.P[p + '.weight'], self.P[p + '.bias'], self.B[p + '.running_mean'], self.B[p + '.running_var'])

    # ------------------------------------------------------------------ forward
    def forward(self, frames, crop, unit_input=False, flip=False, dropout_p=0.5, seed=None):
        """frames (B,T,3,H,W) u8 | f32 device tensor.  Returns (logits [B*T,K] fp32, displ [B*T] | None) and keeps the tape."""
        cfg, P = self.cfg, self.P
        b, t = frames.shape[:2]
        n = b * t
        fr = frames.reshape(n, *frames.shape[2:])
        tape = {'frames': fr, 'crop': crop, 'unit': unit_input, 'flip': flip, 'b': b, 't': t, 'blocks': []}
        y0 = T.stem_raw(fr, unit_input, crop, flip, P['_features.stem.conv.weight'], self.adt)
        st0 = self._bn(y0.view(-1, 32), 32, '_features.stem.bn')
        x = T.bn_act_fwd(y0, st0, relu=True)
        tape['stem'] = (y0, st0, x)
        for p, cin, cout, stride, shifted in cfg.blocks():
            x = self._block_fwd(x, p, cin, cout, stride, shifted, b, t, tape)
        tape['last'] = x
        nfr, h, w, d = x.shape
        feat = ops.pool_posenc(x, t, P['temp_enc'])                      # [n, d] fp32 (+ temp_enc)
        out = self._temporal_fwd(feat.view(b, t, d), tape)
        # heads (Dropout(p) in front of each FC head, model/modules.py:366-376)
        o2 = out.reshape(n, d)
        hd = {}
        if dropout_p > 0:
            if seed is not None:
                self.seed_dev.fill_(int(seed))
            self.seed_dev.add_(1)            # device-side step counter: a captured graph draws fresh masks on every replay
            xin, hd['mask_c'] = T.dropout_fwd_devseed(o2, dropout_p, self.seed_dev, 0)
        else:
            xin = o2
        hd['x_c'] = xin
        if cfg.double_head:              # FC2Layers: cat(fc1(drop(x)), fc2(drop(x))), one Dropout per head (model/modules.py:378-387)
            n1, n2 = cfg.double_head
            if dropout_p > 0:
                x2, hd['mask_c2'] = T.dropout_fwd_devseed(o2, dropout_p, self.seed_dev, 2)
            else:
                x2 = o2
            hd['x_c2'] = x2
            logits = torch.empty((n, n1 + n2), dtype=torch.float32, device=o2.device)
            T.linear_fwd(xin, P['_pred_fine._fc1._fc_out.weight'], P['_pred_fine._fc1._fc_out.bias'], out=logits, col0=0)
            T.linear_fwd(x2, P['_pred_fine._fc2._fc_out.weight'], P['_pred_fine._fc2._fc_out.bias'], out=logits, col0=n1)
        else:
            logits = T.linear_fwd(xin, P['_pred_fine._fc_out.weight'], P['_pred_fine._fc_out.bias'])
        displ = None
        if cfg.radi_displacement > 0:
            if dropout_p > 0:
                xd, hd['mask_d'] = T.dropout_fwd_devseed(o2, dropout_p, self.seed_dev, 1)
            else:
                xd = o2
            hd['x_d'] = xd
            displ = T.linear_fwd(xd, P['_pred_displ._fc_out.weight'], P['_pred_displ._fc_out.bias']).view(-1)
        hd['p'] = dropout_p
        tape['heads'] = hd
        self.tape = tape
        return logits, displ

    def _block_fwd(self, x, p, cin, cout, stride, shifted, clips, clip_len, tape):
        P = self.P
        n, h, w, _ = x.shape
        M = n * h * w
        rec = dict(p=p, cin=cin, cout=cout, stride=stride, shifted=shifted, x=x)
        if shifted:
            fd = fold_dim(cin)
            g = p + '.conv1.gs'
            stg = self._bn(x.view(M, cin), fd, g + '.bn')
            mode = L.SHIFT_GSF if self.cfg.shift_mode == 'gsf' else L.SHIFT_GSM
            w3d = P[g + '.conv3D.weight'].reshape(-1)
            if mode == L.SHIFT_GSF:
                cc_w = torch.cat([P[g + '.channel_conv1.weight'].reshape(-1), P[g + '.channel_conv2.weight'].reshape(-1)])
                cc_b = torch.cat([P[g + '.channel_conv1.bias'], P[g + '.channel_conv2.bias']])
            else:
                cc_w = cc_b = None
            a1, gws = T.gsf_cat_fwd(x, clips, clip_len, fd, mode, stg, w3d, P[g + '.conv3D.bias'], cc_w, cc_b)
            rec.update(fold=fd, stg=stg, gws=gws, mode=mode, w3d=w3d, cc_w=cc_w, cat=a1)
            c1 = p + '.conv1.net'
        else:
            a1 = x.view(M, cin)
            c1 = p + '.conv1'
        rec['c1'] = c1
        w1, w1t = self._wop(c1 + '.conv.weight', cout, cin)
        y1 = self._gemm(a1, w1, M)
        st1 = self._bn(y1, cout, c1 + '.bn')
        z1 = T.bn_act_fwd(y1, st1, relu=True)
        if self.adt == torch.bfloat16:      # tcgen05 implicit GEMM (conv3x3g_tc.cu), weights re-imaged from the fp32 master
            y2 = T.conv3x3g_tc_raw(z1.view(n, h, w, cout), T.conv3_weight_image(P[p + '.conv2.conv.weight'], self.gw), stride)
        else:
            y2 = T.conv3x3g_raw(z1.view(n, h, w, cout), P[p + '.conv2.conv.weight'], self.gw, stride)
        oh, ow = y2.shape[1:3]
        M2 = n * oh * ow
        st2 = self._bn(y2.view(M2, cout), cout, p + '.conv2.bn')
        z2 = T.bn_act_fwd(y2, st2, relu=True)
        rd = P[p + '.se.fc1.weight'].shape[0]
        se_w1 = P[p + '.se.fc1.weight'].reshape(rd, cout)
        se_w2t = P[p + '.se.fc2.weight'].reshape(cout, rd).t().contiguous()
        u, se_ws = T.se_train_fwd(z2, se_w1, P[p + '.se.fc1.bias'], se_w2t, P[p + '.se.fc2.bias'])
        w3, w3t = self._wop(p + '.conv3.conv.weight', cout, cout)
        y3 = self._gemm(u.view(M2, cout), w3, M2)
        st3 = self._bn(y3, cout, p + '.conv3.bn')
        has_ds = (p + '.downsample.conv.weight') in P
        if has_ds:
            wd, wdt = self._wop(p + '.downsample.conv.weight', cout, cin)
            ysc = self._gemm(x.view(M, cin), wd, M2, gather=(stride, h, w) if stride > 1 else None)
            std = self._bn(ysc, cout, p + '.downsample.bn')
            sc = T.bn_act_fwd(ysc, std, relu=False)
            rec.update(wdt=wdt, ysc=ysc, std=std)
        else:
            sc = x.view(M2, cout)
        z3 = T.bn_act_fwd(y3, st3, residual=sc, relu=True).view(n, oh, ow, cout)
        rec.update(w1t=w1t, y1=y1, st1=st1, z1=z1, y2=y2, st2=st2, z2=z2, se_w1=se_w1, se_w2t=se_w2t, se_ws=se_ws, u=u, w3t=w3t,
                   y3=y3, st3=st3, z3=z3, has_ds=has_ds)
        tape['blocks'].append(rec)
        return z3

    # ---- temporal ----
    def _mix_weights(self, p, names):
        d = self.cfg.feat_dim
        out = {}
        for short, key in names:
            out[short + '_w'] = self.P[p + '.' + key + '.weight'].reshape(d, -1)
            out[short + '_b'] = self.P[p + '.' + key + '.bias']
        return out

    def _mlp_fwd(self, g, y2d, p, rows, rec):
        d = self.cfg.feat_dim
        w1, w1t = self._wop(p + '.mlp.0.weight', 4 * d, d)
        w2, w2t = self._wop(p + '.mlp.2.weight', d, 4 * d)
        h = self._gemm(g.view(rows, d), w1, rows, bias=self.P[p + '.mlp.0.bias'], out_dtype=torch.float32)     # pre-GELU
        a = T.gelu_fwd(h, self.adt)
        out = self._gemm(a, w2, rows, bias=self.P[p + '.mlp.2.bias'], residual=y2d, out_dtype=torch.float32)
        rec.update(g=g, h=h, a=a, w1t=w1t, w2t=w2t)
        return out

    def _sgp_fwd(self, x, t_out, idx, tape):
        cfg = self.cfg
        p = '_temp_fine._sgp.%d' % idx
        b, t_in, d = x.shape
        w = self._mix_weights(p, (('psi', 'psi'), ('fc', 'fc'), ('convw', 'convw'), ('convkw', 'convkw'), ('gfc', 'global_fc')))
        w.update(ln_w=self.P[p + '.ln.weight'].reshape(d), ln_b=self.P[p + '.ln.bias'].reshape(d),
                 gn_w=self.P[p + '.gn.weight'], gn_b=self.P[p + '.gn.bias'])
        y, g = ops.sgp_mix(x, t_out, cfg.sgp_ks, cfg.sgp_up, w, self.adt)
        rec = dict(kind='sgp', p=p, x=x, t_out=t_out, y=y, w=w)
        out = self._mlp_fwd(g, y.view(b * t_out, d), p, b * t_out, rec).view(b, t_out, d)
        tape['temporal'].append(rec)
        return out

    def _temporal_fwd(self, feat, tape):
        cfg = self.cfg
        b, t, d = feat.shape
        Ln = cfg.n_layers
        lens = [math.ceil(t / 2 ** i) for i in range(Ln + 1)]
        tape['temporal'] = []
        tape['lens'] = lens
        x = feat
        skips = []
        for i in range(Ln):
            x = self._sgp_fwd(x, lens[i], i, tape)
            skips.append(x)
        x = self._sgp_fwd(x, lens[Ln], Ln, tape)
        for i in range(Ln):
            j = Ln - 1 - i
            p = '_temp_fine._sgpMixer.%d' % j
            rows = b * lens[j]
            w = self._mix_weights(p, (('psi1', 'psi1'), ('psi2', 'psi2'), ('convw1', 'convw1'), ('convkw1', 'convkw1'),
                                      ('convw2', 'convw2'), ('convkw2', 'convkw2'), ('fc1', 'fc1'), ('gfc1', 'global_fc1'),
                                      ('fc2', 'fc2'), ('gfc2', 'global_fc2')))
            for k in ('ln1', 'ln2'):
                w[k + '_w'] = self.P[p + '.%s.weight' % k].reshape(d)
                w[k + '_b'] = self.P[p + '.%s.bias' % k].reshape(d)
            cat = ops.sgp_mixer_mix(x, skips[j], cfg.sgp_ks, cfg.sgp_up, w, torch.float32)       # [rows, 6d] fp32 (kept for bwd)
            cat_op = cat if self.adt == torch.float32 else T.cast(cat, self.adt)
            wc, wct = self._wop(p + '.concat_fc.weight', d, 6 * d)
            pre = self._gemm(cat_op, wc, rows, bias=self.P[p + '.concat_fc.bias'], out_dtype=torch.float32)
            o = T.gelu_fwd(pre, torch.float32)
            g = ops.groupnorm(o.view(b, lens[j], d), self.P[p + '.gn.weight'], self.P[p + '.gn.bias'], self.adt)
            rec = dict(kind='mixer', p=p, x=x, skip=skips[j], j=j, cat=cat, cat_op=cat_op, wct=wct, pre=pre, o=o, w=w)
            x = self._mlp_fwd(g, o, p, rows, rec).view(b, lens[j], d)
            tape['temporal'].append(rec)
            x = self._sgp_fwd(x, lens[j], Ln + i + 1, tape)
        return x

    # ------------------------------------------------------------------ loss
    def loss(self, logits, displ, target_hard=None, target_soft=None, labelD=None, fg_weight=5, dataset=None):
        """dataset: int32 device tensor [B] in {1, 2} — required for the joint-dataset double head (model/model.py:278-306)."""
        k = logits.shape[1]
        cw = None
        if fg_weight != 1:
            cw = self._class_weights.get((k, fg_weight))      # cached: no host->device copy inside a graph capture
            if cw is None:
                cw = torch.tensor([1.0] + [float(fg_weight)] * (k - 1), dtype=torch.float32).to(logits.device)
                self._class_weights[(k, fg_weight)] = cw
        if self.cfg.double_head:
            if dataset is None:
                raise ValueError("double-head training needs batch['dataset'] (1 | 2 per clip)")
            n1, n2 = self.cfg.double_head
            loss, dlogits, ddispl = T.ce_mse_loss_2heads(logits, self.tape['b'], self.tape['t'], n1, n2, dataset, target_hard,
                                                          target_soft, cw, displ, labelD if displ is not None else None)
        else:
            loss, dlogits, ddispl = T.ce_mse_loss(logits, target_hard, target_soft, cw, displ, labelD if displ is not None else None)
        self.tape['dlogits'], self.tape['ddispl'] = dlogits, ddispl
        return loss

    # ------------------------------------------------------------------ backward
    def backward(self):
        self.backward_temporal()
        self.backward_backbone()

    def backward_temporal(self):
        """Heads + ED-SGP-Mixer backward: afterwards every gradient of `_temp_fine.*` / `_pred_*` is final (they are the
        contiguous tail of the flat gradient buffer and ~80 % of its bytes), so a data-parallel caller can start their
        all-reduce while backward_backbone() still runs (tdeed_b200.parallel.GradReducer)."""
        cfg, P, G, tape = self.cfg, self.P, self.G, self.tape
        b, t = tape['b'], tape['t']
        n = b * t
        d = cfg.feat_dim
        hd = tape['heads']
        dlogits, ddispl = tape['dlogits'], tape['ddispl']
        k = dlogits.shape[1]
        if cfg.double_head:
            n1, n2 = cfg.double_head
            dx = None
            for j, (c0, nj, xk, mk) in enumerate(((0, n1, 'x_c', 'mask_c'), (n1, n2, 'x_c2', 'mask_c2'))):
                pj = '_pred_fine._fc%d._fc_out' % (j + 1)
                dl = dlogits[:, c0:c0 + nj]
                T.gemm_tn(dl, hd[xk], nj, d, n, out=G[pj + '.weight'])
                T.colsum(dl, out=G[pj + '.bias'])
                if hd['p'] > 0:
                    dx = T.dropout_bwd(T.linear_bwd_data(dlogits, P[pj + '.weight'], col0=c0, n=nj), hd[mk], hd['p'], add=dx)
                else:
                    dx = T.linear_bwd_data(dlogits, P[pj + '.weight'], add=dx, col0=c0, n=nj)
        else:
            T.gemm_tn(dlogits, hd['x_c'], k, d, n, out=G['_pred_fine._fc_out.weight'])
            T.colsum(dlogits, out=G['_pred_fine._fc_out.bias'])
            dx = T.linear_bwd_data(dlogits, P['_pred_fine._fc_out.weight'])
            if hd['p'] > 0:
                dx = T.dropout_bwd(dx, hd['mask_c'], hd['p'])
        if ddispl is not None:
            dd2 = ddispl.view(n, 1)
            T.gemm_tn(dd2, hd['x_d'], 1, d, n, out=G['_pred_displ._fc_out.weight'])
            T.colsum(dd2, out=G['_pred_displ._fc_out.bias'])
            if hd['p'] > 0:
                dx = T.dropout_bwd(T.linear_bwd_data(dd2, P['_pred_displ._fc_out.weight']), hd['mask_d'], hd['p'], add=dx)
            else:
                dx = T.linear_bwd_data(dd2, P['_pred_displ._fc_out.weight'], add=dx)
        tape['dfeat'] = self._temporal_bwd(dx.view(b, t, d))

    def backward_backbone(self):
        """pool + temp_enc, the RegNetY blocks in reverse, the stem."""
        cfg, P, G, tape = self.cfg, self.P, self.G, self.tape
        b, t = tape['b'], tape['t']
        n = b * t
        d = cfg.feat_dim
        dfeat = tape['dfeat']
        # pool + temp_enc
        last = tape['last']
        nfr, h, w, _ = last.shape
        dz, dte = T.pool_posenc_bwd(dfeat.reshape(n, d), b, t, h * w, d, self.adt)
        G['temp_enc'].copy_(dte)
        dz = dz.view(nfr, h, w, d)
        for rec in reversed(tape['blocks']):
            dz = self._block_bwd(rec, dz, b, t)
        y0, st0, a0 = tape['stem']
        dy0, dg, db, _ = T.bn_act_bwd(dz, y0, y0, st0)
        G['_features.stem.bn.weight'].copy_(dg)
        G['_features.stem.bn.bias'].copy_(db)
        if self.adt == torch.bfloat16:
            T.stem_bwd_weight_tc(tape['frames'], tape['unit'], tape['crop'], tape['flip'], dy0, G['_features.stem.conv.weight'])
        else:
            T.stem_bwd_weight(tape['frames'], tape['unit'], tape['crop'], tape['flip'], dy0, out=G['_features.stem.conv.weight'])
        self.tape = None

    def _dw(self, dy2d, x2d, name, cout, cin, rows, gather=None):
        T.gemm_tn(dy2d, x2d, cout, cin, rows, lda=dy2d.stride(0), ldb=x2d.stride(0), gather=gather,
                  out=self.G[name].view(cout, cin))

    def _block_bwd(self, r, dz, clips, clip_len):
        """dz: gradient w.r.t. the block output z3 (NHWC).  Returns the gradient w.r.t. the block input."""
        P, G = self.P, self.G
        p, cin, cout, stride = r['p'], r['cin'], r['cout'], r['stride']
        x = r['x']
        n, h, w, _ = x.shape
        M = n * h * w
        oh, ow = r['z3'].shape[1:3]
        M2 = n * oh * ow
        # z3 = relu(bn3(y3) + sc)
        dy3, dg, db, g_sc = T.bn_act_bwd(dz.view(M2, cout), r['z3'].view(M2, cout), r['y3'], r['st3'], want_dres=True)
        G[p + '.conv3.bn.weight'].copy_(dg)
        G[p + '.conv3.bn.bias'].copy_(db)
        self._dw(dy3, r['u'].view(M2, cout), p + '.conv3.conv.weight', cout, cout, M2)
        du = self._gemm(dy3, r['w3t'], M2).view(n, oh, ow, cout)
        # squeeze-excite
        dz2, d_w1, d_b1, d_w2, d_b2 = T.se_bwd(r['z2'], du, r['se_w1'], P[p + '.se.fc1.bias'], r['se_w2t'], r['se_ws'])
        G[p + '.se.fc1.weight'].view(d_w1.shape).copy_(d_w1)
        G[p + '.se.fc1.bias'].copy_(d_b1)
        G[p + '.se.fc2.weight'].view(d_w2.shape).copy_(d_w2)
        G[p + '.se.fc2.bias'].copy_(d_b2)
        # conv2 (+BN+ReLU)
        dy2, dg, db, _ = T.bn_act_bwd(dz2, r['y2'], r['y2'], r['st2'])          # mask recomputed from y (no residual)
        G[p + '.conv2.bn.weight'].copy_(dg)
        G[p + '.conv2.bn.bias'].copy_(db)
        z1 = r['z1'].view(n, h, w, cout)
        T.conv3x3g_bwd_weight(z1, dy2, self.gw, stride, out=G[p + '.conv2.conv.weight'])
        if self.adt == torch.bfloat16 and stride == 1:   # data gradient = the same tcgen05 conv with the transposed, flipped kernel
            dz1 = T.conv3x3g_tc_raw(dy2, T.conv3_weight_image(P[p + '.conv2.conv.weight'], self.gw, transpose_flip=True), 1)
        elif self.adt == torch.bfloat16 and stride == 2:  # four parity convolutions over the dy grid, scattered into dx
            dz1 = T.conv3x3g_tc_bwd_data_s2(dy2, (n, h, w, cout), P[p + '.conv2.conv.weight'], self.gw)
        else:
            dz1 = T.conv3x3g_bwd_data(dy2, (n, h, w, cout), P[p + '.conv2.conv.weight'], self.gw, stride)
        # conv1 (+BN+ReLU)
        dy1, dg, db, _ = T.bn_act_bwd(dz1.view(M, cout), r['y1'], r['y1'], r['st1'])
        c1 = r['c1']
        G[c1 + '.bn.weight'].copy_(dg)
        G[c1 + '.bn.bias'].copy_(db)
        a1 = r['cat'] if r['shifted'] else x.view(M, cin)
        self._dw(dy1, a1, c1 + '.conv.weight', cout, cin, M)
        # shortcut
        if r['has_ds']:
            dysc, dg, db, _ = T.bn_act_bwd(g_sc, None, r['ysc'], r['std'])
            G[p + '.downsample.bn.weight'].copy_(dg)
            G[p + '.downsample.bn.bias'].copy_(db)
            if stride > 1 and self.adt == torch.bfloat16:      # compact the strided pixels once, then the dense tcgen05 dW GEMM
                self._dw(dysc, T.strided_gather(x, stride).view(M2, cin), p + '.downsample.conv.weight', cout, cin, M2)
            else:
                self._dw(dysc, x.view(M, cin), p + '.downsample.conv.weight', cout, cin, M2, gather=(stride, h, w) if stride > 1 else None)
            dsc = self._gemm(dysc, r['wdt'], M2)                 # [M2, cin] gradient at the (strided) shortcut pixels
            add = None
        else:
            dsc = None
            add = g_sc                                           # identity shortcut: gradient passes through
        if r['shifted']:
            dcat = self._gemm(dy1, r['w1t'], M)
            g = p + '.conv1.gs'
            dxin, dw3, db3, dcc, dgam, dbet = T.gsf_bwd(x, dcat, add, clips, clip_len, r['fold'], r['mode'], r['stg'], r['w3d'],
                                                       r['cc_w'], r['gws'])
            G[g + '.conv3D.weight'].view(-1).copy_(dw3)
            G[g + '.conv3D.bias'].copy_(db3)
            G[g + '.bn.weight'].copy_(dgam)
            G[g + '.bn.bias'].copy_(dbet)
            if dcc is not None:
                for j in (0, 1):
                    G[g + '.channel_conv%d.weight' % (j + 1)].view(-1).copy_(dcc[j, :18])
                    G[g + '.channel_conv%d.bias' % (j + 1)].copy_(dcc[j, 18:])
        else:
            dxin = self._gemm(dy1, r['w1t'], M, residual=add).view(n, h, w, cin)
        if dsc is not None:
            if stride > 1:
                T.strided_add_(dxin, dsc.view(n, oh, ow, cin), stride)
            else:
                T.strided_add_(dxin, dsc.view(n, h, w, cin), 1)
        return dxin

    # ---- temporal backward ----
    def _mlp_bwd(self, r, dout2d, rows):
        """out = y + W2 gelu(W1 g + b1) + b2.  Returns d g (fp32 [rows, d]); the residual part of d y is dout itself."""
        G = self.G
        p = r['p']
        d = self.cfg.feat_dim
        dout_op = dout2d if self.adt == torch.float32 else T.cast(dout2d, self.adt)
        T.gemm_tn(dout_op, r['a'], d, 4 * d, rows, out=G[p + '.mlp.2.weight'].view(d, 4 * d))
        T.colsum(dout2d, out=G[p + '.mlp.2.bias'])
        da = self._gemm(dout_op, r['w2t'], rows, out_dtype=torch.float32)           # [rows, 4d]
        dh = T.gelu_bwd(r['h'], da, self.adt)
        T.gemm_tn(dh, r['g'].view(rows, d), 4 * d, d, rows, out=G[p + '.mlp.0.weight'].view(4 * d, d))
        T.colsum(dh, out=G[p + '.mlp.0.bias'])
        return self._gemm(dh, r['w1t'], rows, out_dtype=torch.float32)

    def _branch_grads(self, p, pairs):
        return {short: self.G[p + '.' + key].view(self.cfg.feat_dim, -1) if short.endswith('_w') else self.G[p + '.' + key]
                for short, key in pairs}

    def _sgp_bwd(self, r, dout, add_to_input=None):
        """dout [B, t_out, d] -> gradient w.r.t. the block input x [B, t_in, d] (+ add_to_input)."""
        cfg, G = self.cfg, self.G
        p = r['p']
        x, t_out, y, w = r['x'], r['t_out'], r['y'], r['w']
        b, t_in, d = x.shape
        rows = b * t_out
        dg = self._mlp_bwd(r, dout.reshape(rows, d), rows)
        dy, dgw, dgb = T.groupnorm_bwd(y, dg.view(b, t_out, d), w['gn_w'], add=dout.reshape(b, t_out, d))
        G[p + '.gn.weight'].copy_(dgw)
        G[p + '.gn.bias'].copy_(dgb)
        ln, stats, xp, arg = T.chan_ln_fwd(x, t_out, w['ln_w'], w['ln_b'], want_pool=True)
        weights = {k: w[k] for k in T.BRANCH_ORDER}
        grads = self._branch_grads(p, (('psi_w', 'psi.weight'), ('psi_b', 'psi.bias'), ('convw_w', 'convw.weight'),
                                       ('convw_b', 'convw.bias'), ('convkw_w', 'convkw.weight'), ('convkw_b', 'convkw.bias'),
                                       ('fc_w', 'fc.weight'), ('fc_b', 'fc.bias'), ('gfc_w', 'global_fc.weight'),
                                       ('gfc_b', 'global_fc.bias')))
        d_ln = T.sgp_branch_bwd(ln, d, dy, dy, dy, d, b, t_out, d, cfg.sgp_ks, cfg.sgp_up, weights, grads)
        dxp, dlw, dlb = T.chan_ln_bwd(xp.view(rows, d), stats, d_ln, d, w['ln_w'], add=dy.view(rows, d))
        G[p + '.ln.weight'].view(-1).copy_(dlw)
        G[p + '.ln.bias'].view(-1).copy_(dlb)
        if t_in == t_out and add_to_input is None:
            return dxp.view(b, t_in, d)
        return T.maxpool_bwd(dxp.view(b, t_out, d), arg, t_in, add=add_to_input)

    def _mixer_bwd(self, r, dout):
        """dout [B, T, d] -> (d x_coarse [B, tc, d], d skip [B, T, d])."""
        cfg, G = self.cfg, self.G
        p, w = r['p'], r['w']
        xc, skip = r['x'], r['skip']
        b, tc, d = xc.shape
        tl = skip.shape[1]
        rows = b * tl
        dg = self._mlp_bwd(r, dout.reshape(rows, d), rows)
        do, dgw, dgb = T.groupnorm_bwd(r['o'].view(b, tl, d), dg.view(b, tl, d), self.P[p + '.gn.weight'], add=dout.reshape(b, tl, d))
        G[p + '.gn.weight'].copy_(dgw)
        G[p + '.gn.bias'].copy_(dgb)
        dpre = T.gelu_bwd(r['pre'], do.view(rows, d), torch.float32)
        dpre_op = dpre if self.adt == torch.float32 else T.cast(dpre, self.adt)
        T.gemm_tn(dpre_op, r['cat_op'], d, 6 * d, rows, out=G[p + '.concat_fc.weight'].view(d, 6 * d))
        T.colsum(dpre, out=G[p + '.concat_fc.bias'])
        dcat = self._gemm(dpre_op, r['wct'], rows, out_dtype=torch.float32)        # [rows, 6d]
        cat = r['cat']

        def col(tensor, i):
            return tensor.view(-1)[i * d:]          # pointer to column block i (leading dim stays 6d)

        outs = []
        for sfx, ci_conv, ci_fc, ci_id, src, t_src, lnk in (('1', 0, 2, 4, skip, tl, 'ln1'), ('2', 1, 3, 5, xc, tc, 'ln2')):
            weights = {'psi_w': w['psi%s_w' % sfx], 'psi_b': w['psi%s_b' % sfx], 'convw_w': w['convw%s_w' % sfx],
                       'convw_b': w['convw%s_b' % sfx], 'convkw_w': w['convkw%s_w' % sfx], 'convkw_b': w['convkw%s_b' % sfx],
                       'fc_w': w['fc%s_w' % sfx], 'fc_b': w['fc%s_b' % sfx], 'gfc_w': w['gfc%s_w' % sfx], 'gfc_b': w['gfc%s_b' % sfx]}
            grads = self._branch_grads(p, (('psi_w', 'psi%s.weight' % sfx), ('psi_b', 'psi%s.bias' % sfx),
                                           ('convw_w', 'convw%s.weight' % sfx), ('convw_b', 'convw%s.bias' % sfx),
                                           ('convkw_w', 'convkw%s.weight' % sfx), ('convkw_b', 'convkw%s.bias' % sfx),
                                           ('fc_w', 'fc%s.weight' % sfx), ('fc_b', 'fc%s.bias' % sfx),
                                           ('gfc_w', 'global_fc%s.weight' % sfx), ('gfc_b', 'global_fc%s.bias' % sfx)))
            d_lnout = T.sgp_branch_bwd(col(cat, ci_id), 6 * d, col(dcat, ci_conv), col(dcat, ci_fc), col(dcat, ci_id), 6 * d,
                                       b, tl, d, cfg.sgp_ks, cfg.sgp_up, weights, grads)
            if sfx == '2':
                d_lnout = T.upsample_bwd(d_lnout, tc)
            _, stats, _, _ = T.chan_ln_fwd(src, t_src, w[lnk + '_w'], w[lnk + '_b'])
            dsrc, dlw, dlb = T.chan_ln_bwd(src.reshape(b * t_src, d), stats, d_lnout, d, w[lnk + '_w'])
            G[p + '.%s.weight' % lnk].view(-1).copy_(dlw)
            G[p + '.%s.bias' % lnk].view(-1).copy_(dlb)
            outs.append(dsrc.view(b, t_src, d))
        return outs[1], outs[0]

    def _temporal_bwd(self, dout):
        cfg = self.cfg
        Ln = cfg.n_layers
        recs = self.tape['temporal']
        # forward order: sgp 0..L-1 (encoder), sgp L, then for i in 0..L-1: mixer j=L-1-i, sgp L+i+1
        dskip = [None] * Ln
        idx = len(recs) - 1
        d = dout
        for i in reversed(range(Ln)):
            d = self._sgp_bwd(recs[idx], d)
            idx -= 1
            d, ds = self._mixer_bwd(recs[idx], d)
            idx -= 1
            dskip[recs[idx + 1]['j']] = ds
        d = self._sgp_bwd(recs[idx], d, add_to_input=dskip[Ln - 1] if Ln > 0 else None)    # bottleneck block: input = pool(x_{L-1})
        idx -= 1
        for i in reversed(range(Ln)):
            add = dskip[i - 1] if i > 0 else None
            d = self._sgp_bwd(recs[idx], d, add_to_input=add)
            idx -= 1
        return d
