"""GPU parity of every training kernel (include/tdeed_b200_train.h) against PyTorch autograd of the same op in fp32
(TF32 off) — for the composite ops (gate-shift, SGP branches) against the oracle's functions differentiated by autograd.
Tolerances: fp32 storage <= 2e-4 relative-to-max (weight gradients are long fp32 sums); bf16 storage <= 3e-2."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import tdeed_oracle as O


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def T():
    from tdeed_b200 import train_ops
    return train_ops


def LIB():
    from tdeed_b200 import _lib
    return _lib


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


DEV = 'cuda'


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize('C,M', [(24, 5000), (368, 777), (32, 40000)])
def test_bn_train_fwd_bwd(dtype, tol, C, M):
    g = torch.Generator(device=DEV).manual_seed(C + M)
    y = (torch.randn((M, C), device=DEV, generator=g) * 2 + 3).to(dtype)
    res = torch.randn((M, C), device=DEV, generator=g).to(dtype)
    gamma = torch.rand(C, device=DEV, generator=g) + 0.5
    beta = torch.randn(C, device=DEV, generator=g) * 0.1
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    dz = torch.randn((M, C), device=DEV, generator=g).to(dtype)
    # reference
    yr = y.float().requires_grad_(True)
    rr = res.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()
    zr = F.relu(F.batch_norm(yr, rm_r, rv_r, gr, br, training=True, momentum=0.1, eps=1e-5) + rr)
    zr.backward(dz.float())
    # ours
    stats = T().bn_stats(y, C, gamma, beta, rm, rv)
    z = T().bn_act_fwd(y, stats, residual=res, relu=True)
    assert rel(z.float(), zr) < tol
    assert rel(rm, rm_r) < 1e-5 and rel(rv, rv_r) < 1e-4
    if dtype == torch.float32:
        dy, dgamma, dbeta, dres = T().bn_act_bwd(dz, z, y, stats, want_dres=True)
        assert rel(dy, yr.grad) < 1e-4
        assert rel(dres, rr.grad) < 1e-6
        assert rel(dgamma, gr.grad) < 1e-4 and rel(dbeta, br.grad) < 1e-4
    else:   # the ReLU mask comes from the bf16-rounded z: compare against a reference that uses the same mask
        dy, dgamma, dbeta, dres = T().bn_act_bwd(dz, z, y, stats, want_dres=True)
        mask = (z.float() > 0).float()
        gmask = dz.float() * mask
        xhat = (y.float() - stats[0]) * stats[1]
        dg_ref = (gmask * xhat).sum(0)
        db_ref = gmask.sum(0)
        dy_ref = stats[2] * (gmask - db_ref / M - xhat * dg_ref / M)
        assert rel(dy.float(), dy_ref) < tol and rel(dgamma, dg_ref) < 1e-3 and rel(dbeta, db_ref) < 1e-3


def test_bn_stats_column_slice_and_no_relu():
    g = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn((3000, 368), device=DEV, generator=g) * 1.5 + 0.7
    fold = 92
    gamma = torch.rand(fold, device=DEV, generator=g) + 0.5
    beta = torch.randn(fold, device=DEV, generator=g)
    stats = T().bn_stats(x, fold, gamma, beta)
    xs = x[:, :fold]
    assert rel(stats[0], xs.mean(0)) < 1e-5
    assert rel(stats[1], 1 / torch.sqrt(xs.var(0, unbiased=False) + 1e-5)) < 1e-5
    # no-ReLU backward (conv3 / downsample BN)
    y = torch.randn((1000, 56), device=DEV, generator=g)
    ga, be = torch.rand(56, device=DEV, generator=g) + 0.5, torch.zeros(56, device=DEV)
    dz = torch.randn((1000, 56), device=DEV, generator=g)
    yr, gr = y.clone().requires_grad_(True), ga.clone().requires_grad_(True)
    F.batch_norm(yr, None, None, gr, be, training=True).backward(dz)
    st = T().bn_stats(y, 56, ga, be)
    dy, dgamma, dbeta, _ = T().bn_act_bwd(dz, None, y, st)
    assert rel(dy, yr.grad) < 1e-4 and rel(dgamma, gr.grad) < 1e-4
    # ReLU without residual: passing z == y recomputes the mask from y instead of reading z
    yr2, gr2 = y.clone().requires_grad_(True), ga.clone().requires_grad_(True)
    F.relu(F.batch_norm(yr2, None, None, gr2, be, training=True)).backward(dz)
    dy2, dgamma2, _, _ = T().bn_act_bwd(dz, y, y, st)
    assert rel(dy2, yr2.grad) < 1e-4 and rel(dgamma2, gr2.grad) < 1e-4


@pytest.mark.parametrize('dt_a,dt_b', [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)])
@pytest.mark.parametrize('R,m,n', [(5000, 24, 32), (70000, 152, 152), (300, 5, 368), (1000, 368, 92)])
def test_gemm_tn(dt_a, dt_b, R, m, n):
    g = torch.Generator(device=DEV).manual_seed(R + m)
    a = torch.randn((R, m), device=DEV, generator=g).to(dt_a)
    b = torch.randn((R, n), device=DEV, generator=g).to(dt_b)
    out = T().gemm_tn(a, b, m, n, R)
    ref = a.double().t() @ b.double()
    assert rel(out, ref) < 2e-5


@pytest.mark.parametrize('R,m,n', [(4096, 64, 64), (20000, 320, 320), (3000, 768, 3072), (777, 152, 88), (100000, 24, 32),
                                   (6400, 3072, 768), (1500, 368, 2208)])
def test_gemm_tn_tcgen05(R, m, n):
    """bf16 x bf16 with 8-aligned leading dims takes the tcgen05 MN-major kernel (train_gemm_tc.cu)."""
    g = torch.Generator(device=DEV).manual_seed(R + m + n)
    a = torch.randn((R, m), device=DEV, generator=g).bfloat16()
    b = torch.randn((R, n), device=DEV, generator=g).bfloat16()
    out = T().gemm_tn(a, b, m, n, R)
    ref = a.double().t() @ b.double()
    assert rel(out, ref) < 1e-4
    # column slices of wider tensors (leading dim > columns), scaled
    wide_a = torch.randn((R, m + 16), device=DEV, generator=g).bfloat16()
    out2 = T().gemm_tn(wide_a[:, 8:], b, m, n, R, alpha=0.5)
    assert rel(out2, 0.5 * (wide_a[:, 8:8 + m].double().t() @ b.double())) < 1e-4


def test_gemm_tn_gather_colsum_strided_add():
    g = torch.Generator(device=DEV).manual_seed(3)
    nfr, h, w, cin, cout = 6, 9, 11, 24, 56
    x = torch.randn((nfr, h, w, cin), device=DEV, generator=g)
    oh, ow = (h + 1) // 2, (w + 1) // 2
    dy = torch.randn((nfr * oh * ow, cout), device=DEV, generator=g)
    out = T().gemm_tn(dy, x, cout, cin, nfr * oh * ow, ldb=cin, gather=(2, h, w))
    xs = x[:, ::2, ::2].reshape(-1, cin)
    assert rel(out, dy.double().t() @ xs.double()) < 2e-5
    assert rel(T().colsum(dy), dy.double().sum(0)) < 1e-5
    dst = torch.randn((nfr, h, w, cin), device=DEV, generator=g)
    src = torch.randn((nfr, oh, ow, cin), device=DEV, generator=g)
    ref = dst.clone()
    ref[:, ::2, ::2] += src
    T().strided_add_(dst, src, 2)
    assert torch.equal(dst, ref)
    assert torch.equal(T().strided_gather(x, 2), x[:, ::2, ::2].contiguous())
    assert torch.equal(T().strided_gather(x.bfloat16(), 2), x.bfloat16()[:, ::2, ::2].contiguous())


@pytest.mark.parametrize('unit', [False, True])
def test_stem_raw_and_weight_grad(unit):
    g = torch.Generator(device=DEV).manual_seed(7)
    n, H, W = 5, 70, 90
    crop = (3, 5, 64, 80)
    frames = torch.randint(0, 256, (n, 3, H, W), device=DEV, generator=g).float()
    if unit:
        frames = frames / 255.
    wgt = torch.randn((32, 3, 3, 3), device=DEV, generator=g) * 0.2
    cy, cx, h, w = crop
    x = frames[:, :, cy:cy + h, cx:cx + w]
    x = x if unit else x / 255.
    mean = torch.tensor(O.IMAGENET_MEAN, device=DEV).view(1, 3, 1, 1)
    std = torch.tensor(O.IMAGENET_STD, device=DEV).view(1, 3, 1, 1)
    wr = wgt.clone().requires_grad_(True)
    yr = F.conv2d((x - mean) / std, wr, stride=2, padding=1)
    dy = torch.randn(yr.shape, device=DEV, generator=g)
    yr.backward(dy)
    y = T().stem_raw(frames, unit, crop, False, wgt, torch.float32)
    assert rel(nchw(y), yr) < 1e-5
    dw = T().stem_bwd_weight(frames, unit, crop, False, nhwc(dy))
    assert rel(dw, wr.grad) < 1e-4
    # bf16 training path: im2col (bf16 patches) + tcgen05 dW GEMM
    dwt = T().stem_bwd_weight_tc(frames, unit, crop, False, nhwc(dy).bfloat16(), torch.empty((32, 3, 3, 3), device=DEV))
    assert rel(dwt, wr.grad) < 1e-2


@pytest.mark.parametrize('c,gw,stride,h,w', [(24, 8, 2, 32, 32), (56, 8, 1, 17, 23), (64, 16, 2, 30, 45), (320, 16, 1, 7, 7), (368, 8, 1, 7, 7),
                                              (128, 16, 2, 13, 9)])
def test_conv3x3g_train(c, gw, stride, h, w):
    g = torch.Generator(device=DEV).manual_seed(c + h)
    n = 6
    x = torch.randn((n, c, h, w), device=DEV, generator=g)
    wgt = torch.randn((c, gw, 3, 3), device=DEV, generator=g) * 0.2
    xr, wr = x.clone().requires_grad_(True), wgt.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=stride, padding=1, groups=c // gw)
    dy = torch.randn(yr.shape, device=DEV, generator=g)
    yr.backward(dy)
    y = T().conv3x3g_raw(nhwc(x), wgt, gw, stride)
    assert rel(nchw(y), yr) < 1e-5
    dx = T().conv3x3g_bwd_data(nhwc(dy), (n, h, w, c), wgt, gw, stride)
    assert rel(nchw(dx), xr.grad) < 1e-5
    dw = T().conv3x3g_bwd_weight(nhwc(x), nhwc(dy), gw, stride)
    assert rel(dw, wr.grad) < 1e-4
    # bf16 storage
    dxb = T().conv3x3g_bwd_data(nhwc(dy).bfloat16(), (n, h, w, c), wgt, gw, stride)
    assert rel(nchw(dxb.float()), xr.grad) < 2e-2
    dwb = T().conv3x3g_bwd_weight(nhwc(x).bfloat16(), nhwc(dy).bfloat16(), gw, stride)
    assert rel(dwb, wr.grad) < 2e-2
    # the bf16 path is the tcgen05 kernel (train_conv_tc.cu): exact products of the rounded inputs, fp32 accumulation
    xq, wq = x.bfloat16().float().requires_grad_(True), wgt.clone().requires_grad_(True)
    F.conv2d(xq, wq, stride=stride, padding=1, groups=c // gw).backward(dy.bfloat16().float())
    assert rel(dwb, wq.grad) < 2e-4
    # tcgen05 raw forward / stride-1 data gradient with device-built weight images (weights rounded to bf16 inside)
    wb = wgt.bfloat16().float()
    yq = F.conv2d(x.bfloat16().float(), wb, stride=stride, padding=1, groups=c // gw)
    yt = T().conv3x3g_tc_raw(nhwc(x).bfloat16(), T().conv3_weight_image(wgt, gw), stride)
    assert rel(nchw(yt.float()), yq) < 1e-2
    if stride == 1:
        dq = dy.bfloat16().float()
        dxq = torch.nn.grad.conv2d_input(x.shape, wb, dq, stride=1, padding=1, groups=c // gw)
        dxt = T().conv3x3g_tc_raw(nhwc(dy).bfloat16(), T().conv3_weight_image(wgt, gw, transpose_flip=True), 1)
        assert rel(nchw(dxt.float()), dxq) < 1e-2
    else:
        dq = dy.bfloat16().float()
        dxq = torch.nn.grad.conv2d_input(x.shape, wb, dq, stride=2, padding=1, groups=c // gw)
        dxt = T().conv3x3g_tc_bwd_data_s2(nhwc(dy).bfloat16(), (n, h, w, c), wgt, gw)
        assert rel(nchw(dxt.float()), dxq) < 1e-2


@pytest.mark.parametrize('c,rd,hw', [(24, 8, 64), (152, 38, 49), (768, 80, 16)])
def test_se_train(c, rd, hw):
    g = torch.Generator(device=DEV).manual_seed(c)
    n = 10
    side = int(math.isqrt(hw))
    x = torch.randn((n, c, side, side), device=DEV, generator=g)
    w1 = torch.randn((rd, c), device=DEV, generator=g) * 0.2
    b1 = torch.randn(rd, device=DEV, generator=g) * 0.1
    w2 = torch.randn((c, rd), device=DEV, generator=g) * 0.2
    b2 = torch.randn(c, device=DEV, generator=g) * 0.1
    ps = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    s = ps[0].mean(dim=(2, 3))
    s = torch.sigmoid(F.linear(F.relu(F.linear(s, ps[1], ps[2])), ps[3], ps[4]))
    ur = ps[0] * s[:, :, None, None]
    du = torch.randn(ur.shape, device=DEV, generator=g)
    ur.backward(du)
    w2t = w2.t().contiguous()
    u, ws = T().se_train_fwd(nhwc(x), w1, b1, w2t, b2)
    assert rel(nchw(u), ur) < 1e-5
    dx, d_w1, d_b1, d_w2, d_b2 = T().se_bwd(nhwc(x), nhwc(du), w1, b1, w2t, ws)
    assert rel(nchw(dx), ps[0].grad) < 1e-4
    assert rel(d_w1, ps[1].grad) < 1e-4 and rel(d_b1, ps[2].grad) < 1e-4
    assert rel(d_w2, ps[3].grad) < 1e-4 and rel(d_b2, ps[4].grad) < 1e-4
    # bf16 storage
    xb, dub = nhwc(x).bfloat16(), nhwc(du).bfloat16()
    ps2 = [t.clone().requires_grad_(True) for t in (nchw(xb.float()), w1, b1, w2, b2)]
    s2 = torch.sigmoid(F.linear(F.relu(F.linear(ps2[0].mean(dim=(2, 3)), ps2[1], ps2[2])), ps2[3], ps2[4]))
    (ps2[0] * s2[:, :, None, None]).backward(nchw(dub.float()))
    ub, wsb = T().se_train_fwd(xb, w1, b1, w2t, b2)
    dxb, d_w1b, _, d_w2b, _ = T().se_bwd(xb, dub, w1, b1, w2t, wsb)
    assert rel(nchw(dxb.float()), ps2[0].grad) < 1e-2
    assert rel(d_w1b, ps2[1].grad) < 1e-3 and rel(d_w2b, ps2[3].grad) < 1e-3


def test_pool_posenc_bwd():
    g = torch.Generator(device=DEV).manual_seed(1)
    clips, clip_len, hw, c = 3, 5, 12, 40
    df = torch.randn((clips * clip_len, c), device=DEV, generator=g)
    dz, dte = T().pool_posenc_bwd(df, clips, clip_len, hw, c, torch.float32)
    assert rel(dz, (df / hw)[:, None, :].expand(-1, hw, -1)) < 1e-6
    assert rel(dte, df.view(clips, clip_len, c).sum(0)) < 1e-6


def _gs_params(fold, mode, g):
    sd = {'gs.bn.weight': torch.rand(fold, device=DEV, generator=g) + 0.5,
          'gs.bn.bias': torch.randn(fold, device=DEV, generator=g) * 0.2,
          'gs.conv3D.weight': torch.randn((2, fold // 2, 3, 3, 3), device=DEV, generator=g) / math.sqrt(fold * 13.5),
          'gs.conv3D.bias': torch.randn(2, device=DEV, generator=g) * 0.1}
    if mode == 'gsf':
        for j in (1, 2):
            sd['gs.channel_conv%d.weight' % j] = torch.randn((1, 2, 3, 3), device=DEV, generator=g) * 0.4
            sd['gs.channel_conv%d.bias' % j] = torch.randn(1, device=DEV, generator=g) * 0.1
    return sd


@pytest.mark.parametrize('mode', ['gsf', 'gsm'])
@pytest.mark.parametrize('c,fold,h,w,clips,clip_len', [(56, 16, 8, 8, 2, 6), (152, 40, 5, 7, 2, 5), (368, 92, 3, 3, 1, 7)])
def test_gate_shift_train_fwd_bwd(mode, c, fold, h, w, clips, clip_len):
    g = torch.Generator(device=DEV).manual_seed(fold + h)
    n = clips * clip_len
    x = torch.randn((n, c, h, w), device=DEV, generator=g) + 0.3
    sd = _gs_params(fold, mode, g)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    cat_r = torch.cat([O.gate_shift(xr[:, :fold], sdr, 'gs', clip_len, mode, train=True), xr[:, fold:]], dim=1)
    dcat = torch.randn(cat_r.shape, device=DEV, generator=g)
    add = torch.randn(cat_r.shape, device=DEV, generator=g)
    cat_r.backward(dcat)
    L = LIB()
    m = L.SHIFT_GSF if mode == 'gsf' else L.SHIFT_GSM
    xh = nhwc(x)
    stats = T().bn_stats(xh.view(-1, c), fold, sd['gs.bn.weight'], sd['gs.bn.bias'])
    if mode == 'gsf':
        cc_w = torch.cat([sd['gs.channel_conv1.weight'].reshape(-1), sd['gs.channel_conv2.weight'].reshape(-1)])
        cc_b = torch.cat([sd['gs.channel_conv1.bias'], sd['gs.channel_conv2.bias']])
    else:
        cc_w = cc_b = None
    w3 = sd['gs.conv3D.weight'].reshape(-1).contiguous()
    cat, ws = T().gsf_cat_fwd(xh, clips, clip_len, fold, m, stats, w3, sd['gs.conv3D.bias'], cc_w, cc_b)
    assert rel(nchw(cat.view(n, h, w, c)), cat_r) < 2e-5
    dx, dw3, db3, dcc, dgam, dbet = T().gsf_bwd(xh, nhwc(dcat), nhwc(add), clips, clip_len, fold, m, stats, w3, cc_w, ws)
    assert rel(nchw(dx), xr.grad + add) < 2e-4
    assert rel(dw3, sdr['gs.conv3D.weight'].grad.reshape(-1)) < 2e-4
    assert rel(db3, sdr['gs.conv3D.bias'].grad) < 2e-4
    assert rel(dgam, sdr['gs.bn.weight'].grad) < 2e-4 and rel(dbet, sdr['gs.bn.bias'].grad) < 2e-4
    if mode == 'gsf':
        for j in (0, 1):
            assert rel(dcc[j, :18], sdr['gs.channel_conv%d.weight' % (j + 1)].grad.reshape(-1)) < 2e-4
            assert rel(dcc[j, 18:], sdr['gs.channel_conv%d.bias' % (j + 1)].grad) < 2e-4
    # bf16 storage of x / d_cat: same kernels, inputs rounded -> compare against autograd on the rounded inputs
    xb, dcb, adb = nhwc(x).bfloat16(), nhwc(dcat).bfloat16(), nhwc(add).bfloat16()
    xr2 = nchw(xb.float()).requires_grad_(True)
    sdr2 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    cat2 = torch.cat([O.gate_shift(xr2[:, :fold], sdr2, 'gs', clip_len, mode, train=True), xr2[:, fold:]], dim=1)
    cat2.backward(nchw(dcb.float()))
    stats_b = T().bn_stats(xb.view(-1, c), fold, sd['gs.bn.weight'], sd['gs.bn.bias'])
    catb, wsb = T().gsf_cat_fwd(xb, clips, clip_len, fold, m, stats_b, w3, sd['gs.conv3D.bias'], cc_w, cc_b)
    assert rel(nchw(catb.view(n, h, w, c).float()), cat2) < 1e-2
    dxb, dw3b, db3b, dccb, dgamb, dbetb = T().gsf_bwd(xb, dcb, adb, clips, clip_len, fold, m, stats_b, w3, cc_w, wsb)
    assert rel(nchw(dxb.float()), xr2.grad + nchw(adb.float())) < 2e-2
    assert rel(dw3b, sdr2['gs.conv3D.weight'].grad.reshape(-1)) < 2e-3
    assert rel(dgamb, sdr2['gs.bn.weight'].grad) < 2e-3


def _branch_sd(C, ks, up, g, sfx=''):
    sd = {}
    for n, k in (('psi', ks), ('fc', 1), ('convw', ks), ('convkw', up), ('global_fc', 1)):
        sd['p.%s%s.weight' % (n, sfx)] = torch.randn((C, 1, k), device=DEV, generator=g) * (0.3 / math.sqrt(k) + 0.05)
        sd['p.%s%s.bias' % (n, sfx)] = torch.randn(C, device=DEV, generator=g) * 0.1
    return sd


@pytest.mark.parametrize('B,t_in,Tn,C,ks,up', [(2, 12, 12, 368, 5, 25), (3, 25, 13, 48, 9, 41), (2, 100, 50, 768, 7, 33)])
def test_sgp_block_backward_pieces(B, t_in, Tn, C, ks, up):
    """LayerNorm(+max-pool) fwd/bwd + branch backward + GroupNorm backward == autograd of the oracle's SGPBlock (without MLP)."""
    g = torch.Generator(device=DEV).manual_seed(C + Tn)
    x = torch.randn((B, t_in, C), device=DEV, generator=g)
    sd = _branch_sd(C, ks, up, g)
    sd['p.ln.weight'] = (torch.rand(C, device=DEV, generator=g) + 0.5).view(1, C, 1)
    sd['p.ln.bias'] = (torch.randn(C, device=DEV, generator=g) * 0.1).view(1, C, 1)
    sd['p.gn.weight'] = torch.rand(C, device=DEV, generator=g) + 0.5
    sd['p.gn.bias'] = torch.randn(C, device=DEV, generator=g) * 0.1
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    xp_r = F.adaptive_max_pool1d(xr.permute(0, 2, 1), Tn) if t_in != Tn else xr.permute(0, 2, 1)
    out = O.channel_layernorm(xp_r, sdr['p.ln.weight'], sdr['p.ln.bias'])
    psi, fc = O._dw(out, sdr, 'p.psi'), O._dw(out, sdr, 'p.fc')
    convw, convkw = O._dw(out, sdr, 'p.convw'), O._dw(out, sdr, 'p.convkw')
    phi = F.relu(O._dw(out.mean(dim=-1, keepdim=True), sdr, 'p.global_fc'))
    y_r = xp_r + fc * phi + (convw + convkw) * psi + out
    g_r = F.group_norm(y_r, 16, sdr['p.gn.weight'], sdr['p.gn.bias'])
    dg = torch.randn((B, Tn, C), device=DEV, generator=g)
    dres = torch.randn((B, Tn, C), device=DEV, generator=g)
    (g_r.permute(0, 2, 1) * dg + y_r.permute(0, 2, 1) * dres).sum().backward()
    # ours
    lnw, lnb = sd['p.ln.weight'].reshape(-1).contiguous(), sd['p.ln.bias'].reshape(-1).contiguous()
    ln, stats, xp, arg = T().chan_ln_fwd(x, Tn, lnw, lnb, want_pool=True)
    assert rel(ln, out.permute(0, 2, 1)) < 1e-5
    y = y_r.detach().permute(0, 2, 1).contiguous()
    dy, dgn_w, dgn_b = T().groupnorm_bwd(y, dg, sd['p.gn.weight'], add=dres)
    assert rel(dgn_w, sdr['p.gn.weight'].grad) < 1e-4 and rel(dgn_b, sdr['p.gn.bias'].grad) < 1e-4
    names = {'psi_w': 'p.psi.weight', 'psi_b': 'p.psi.bias', 'convw_w': 'p.convw.weight', 'convw_b': 'p.convw.bias',
             'convkw_w': 'p.convkw.weight', 'convkw_b': 'p.convkw.bias', 'fc_w': 'p.fc.weight', 'fc_b': 'p.fc.bias',
             'gfc_w': 'p.global_fc.weight', 'gfc_b': 'p.global_fc.bias'}
    weights = {k: sd[v].reshape(sd[v].shape[0], -1).contiguous() for k, v in names.items()}
    grads = {k: torch.empty_like(v) for k, v in weights.items()}
    d_ln = T().sgp_branch_bwd(ln, C, dy, dy, dy, C, B, Tn, C, ks, up, weights, grads)
    for k, v in names.items():
        assert rel(grads[k].reshape(-1), sdr[v].grad.reshape(-1)) < 2e-4, k
    dxp, dlw, dlb = T().chan_ln_bwd(xp.view(-1, C), stats, d_ln, C, lnw, add=dy)
    assert rel(dlw, sdr['p.ln.weight'].grad.reshape(-1)) < 2e-4 and rel(dlb, sdr['p.ln.bias'].grad.reshape(-1)) < 2e-4
    dx = T().maxpool_bwd(dxp.view(B, Tn, C), arg, t_in)
    assert rel(dx, xr.grad) < 2e-4


def test_gelu_upsample_cast():
    g = torch.Generator(device=DEV).manual_seed(2)
    h = torch.randn((50, 64), device=DEV, generator=g) * 2
    hr = h.clone().requires_grad_(True)
    a = F.gelu(hr)
    da = torch.randn(h.shape, device=DEV, generator=g)
    a.backward(da)
    assert rel(T().gelu_fwd(h, torch.float32), a) < 1e-6
    assert rel(T().gelu_bwd(h, da, torch.float32), hr.grad) < 1e-5
    assert rel(T().gelu_fwd(h, torch.bfloat16).float(), a) < 1e-2
    for tc, Tn in ((13, 25), (50, 100), (3, 6)):
        x = torch.randn((2, 24, tc), device=DEV, generator=g, requires_grad=True)
        xu = F.interpolate(x, size=Tn, mode='linear', align_corners=True)
        d = torch.randn(xu.shape, device=DEV, generator=g)
        xu.backward(d)
        dx = T().upsample_bwd(d.permute(0, 2, 1).contiguous(), tc)
        assert rel(dx, x.grad.permute(0, 2, 1)) < 1e-5
    assert torch.equal(T().cast(h, torch.bfloat16), h.bfloat16())


@pytest.mark.parametrize('soft', [False, True])
@pytest.mark.parametrize('K,displ', [(5, True), (33, False)])
def test_loss_and_heads(soft, K, displ):
    g = torch.Generator(device=DEV).manual_seed(K)
    M, C = 300, 368
    feat = torch.randn((M, C), device=DEV, generator=g)
    W = torch.randn((K, C), device=DEV, generator=g) * 0.1
    b = torch.randn(K, device=DEV, generator=g) * 0.1
    Wd = torch.randn((1, C), device=DEV, generator=g) * 0.1
    bd = torch.randn(1, device=DEV, generator=g)
    cw = torch.tensor([1.] + [5.] * (K - 1), device=DEV)
    if soft:
        t = torch.softmax(torch.randn((M, K), device=DEV, generator=g) * 3, dim=1)
        hard, softt = None, t
    else:
        t = torch.randint(0, K, (M,), device=DEV, generator=g)
        hard, softt = t, None
    labelD = torch.randint(-2, 3, (M,), device=DEV, generator=g).float()
    fr = feat.clone().requires_grad_(True)
    Wr, br = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    logits_r = F.linear(fr, Wr, br)
    loss_r = F.cross_entropy(logits_r, t, weight=cw)
    if displ:
        Wdr = Wd.clone().requires_grad_(True)
        d_r = F.linear(fr, Wdr, bd).squeeze(-1)
        loss_r = loss_r + F.mse_loss(d_r, labelD, reduction='none').mean()
    loss_r.backward()
    logits = T().linear_fwd(feat, W, b)
    assert rel(logits, logits_r) < 1e-5
    d = T().linear_fwd(feat, Wd, bd).view(-1) if displ else None
    loss, dlogits, ddispl = T().ce_mse_loss(logits, hard, softt, cw, d, labelD if displ else None)
    assert abs(float(loss[0]) - float(loss_r)) < 1e-5 * max(1.0, abs(float(loss_r)))
    dfeat = T().linear_bwd_data(dlogits, W)
    if displ:
        dfeat = T().linear_bwd_data(ddispl.view(-1, 1), Wd, add=dfeat)
        assert rel(T().gemm_tn(ddispl.view(-1, 1), feat, 1, C, M), Wdr.grad) < 1e-4
    assert rel(dfeat, fr.grad) < 1e-4
    assert rel(T().gemm_tn(dlogits, feat, K, C, M), Wr.grad) < 1e-4
    assert rel(T().colsum(dlogits), br.grad) < 1e-4


def test_mixup_u8_matches_the_reference_arithmetic():
    g = torch.Generator(device=DEV).manual_seed(11)
    a = torch.randint(0, 256, (3, 4, 3, 16, 24), device=DEV, generator=g, dtype=torch.uint8)
    b = torch.randint(0, 256, a.shape, device=DEV, generator=g, dtype=torch.uint8)
    l = torch.tensor([0.3137, 0.9999, 1e-6], dtype=torch.float64)
    lam = torch.stack([l, 1 - l], dim=1).float().to(DEV)
    ref = lam[:, 0].view(3, 1, 1, 1, 1) * a.float() + lam[:, 1].view(3, 1, 1, 1, 1) * b.float()      # model/model.py:240-244
    assert torch.equal(T().mixup_u8(a, b, lam), ref)


def test_loss_flags_out_of_range_labels():
    """F.cross_entropy raises on a target outside [0, K); the fused loss must not read out of bounds and poisons the loss (NaN)."""
    g = torch.Generator(device=DEV).manual_seed(3)
    M, K = 64, 5
    logits = torch.randn((M, K), device=DEV, generator=g)
    cw = torch.ones(K, device=DEV)
    t = torch.randint(0, K, (M,), device=DEV, generator=g)
    loss, dl, _ = T().ce_mse_loss(logits, t, None, cw, None, None)
    assert bool(torch.isfinite(loss).all())
    for bad in (K, -1, 10 ** 9):
        tb = t.clone()
        tb[17] = bad
        loss, dl, _ = T().ce_mse_loss(logits, tb, None, cw, None, None)
        assert bool(torch.isnan(loss[0])) and bool(torch.isnan(loss[1])) and bool(torch.isfinite(dl).all())
    # two heads: a dataset-2 label that was not shifted by n1 (update_labels_2heads skipped) is out of its head's range
    B, Tn, n1, n2 = 2, 16, 3, 4
    lg = torch.randn((B * Tn, n1 + n2), device=DEV, generator=g)
    ds = torch.tensor([1, 2], dtype=torch.int32, device=DEV)
    hard = torch.cat([torch.randint(0, n1, (Tn,), device=DEV, generator=g), n1 + torch.randint(0, n2, (Tn,), device=DEV, generator=g)])
    loss, _, _ = T().ce_mse_loss_2heads(lg, B, Tn, n1, n2, ds, hard, None, torch.ones(n1 + n2, device=DEV), None, None)
    assert bool(torch.isfinite(loss).all())
    hard[Tn + 3] = 1                                      # dataset 2, label < n1
    loss, _, _ = T().ce_mse_loss_2heads(lg, B, Tn, n1, n2, ds, hard, None, torch.ones(n1 + n2, device=DEV), None, None)
    assert bool(torch.isnan(loss[0]))


def test_dropout_and_adamw():
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn((400, 368), device=DEV, generator=g)
    out, mask = T().dropout_fwd(x, 0.5, 1234)
    keep = mask.bool()
    assert 0.45 < keep.float().mean().item() < 0.55
    assert torch.equal(out[keep], x[keep] / 0.5) and float(out[~keep].abs().max()) == 0.0
    out2, mask2 = T().dropout_fwd(x, 0.5, 1235)
    assert not torch.equal(mask, mask2)
    dy = torch.randn(x.shape, device=DEV, generator=g)
    assert torch.equal(T().dropout_bwd(dy, mask, 0.5), torch.where(keep, dy / 0.5, torch.zeros_like(dy)))
    # AdamW: 3 steps against torch.optim.AdamW
    p0 = torch.randn(10007, device=DEV, generator=g)
    grads = [torch.randn(10007, device=DEV, generator=g) * (0.1 + i) for i in range(3)]
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    shadow = torch.empty(p.shape, dtype=torch.bfloat16, device=DEV)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        T().adamw_step_(p, gr, m, v, 1e-3, (0.9, 0.999), 1e-8, 0.01, i + 1, shadow=shadow)
    assert rel(p, pr) < 1e-6
    assert torch.equal(shadow, p.bfloat16())
    y = torch.ones(100, device=DEV)
    T().axpy_(torch.full((100,), 2.0, device=DEV), 0.5, y)
    assert torch.equal(y, torch.full((100,), 2.0, device=DEV))


@pytest.mark.parametrize('in_dtype', [torch.uint8, torch.float32])
def test_fused_augmentation_matches_torchvision(in_dtype):
    """train_aug.cu against torchvision's own tensor ops with the same parameters (model/model.py:77-84 pipeline order)."""
    import torchvision.transforms.functional as TF
    from tdeed_b200.augment import ClipAugment
    g = torch.Generator(device=DEV).manual_seed(3)
    B, Tn, H, W = 6, 4, 40, 52
    crop = (3, 5, 32, 40)
    frames = torch.randint(0, 256, (B, Tn, 3, H, W), device=DEV, generator=g, dtype=torch.uint8)
    # smooth images too (blur / hue behave differently on noise and on gradients)
    yy, xx = torch.meshgrid(torch.arange(H, device=DEV), torch.arange(W, device=DEV), indexing='ij')
    frames[1] = ((yy * 3 + xx * 2) % 256).to(torch.uint8)[None, None].expand(Tn, 3, H, W).clone()
    frames[1, :, 1] = ((yy * 5) % 256).to(torch.uint8)
    if in_dtype == torch.float32:
        frames = frames.float()
    params = [dict(hue=0.13, sat=None, bri=None, con=None, sigma=None, flip=False),
              dict(hue=-0.2, sat=0.8, bri=1.15, con=0.75, sigma=1.3, flip=True),
              dict(hue=None, sat=1.2, bri=None, con=None, sigma=None, flip=True),
              dict(hue=None, sat=None, bri=0.7, con=1.2, sigma=None, flip=False),
              dict(hue=None, sat=None, bri=None, con=None, sigma=0.4, flip=False),
              dict(hue=None, sat=None, bri=None, con=None, sigma=None, flip=False)]
    out = ClipAugment.apply(frames, crop, params)
    cy, cx, h, w = crop
    for i, p in enumerate(params):
        x = frames[i][..., cy:cy + h, cx:cx + w].float() / 255.
        if p['hue'] is not None:
            x = TF.adjust_hue(x, p['hue'])
        if p['sat'] is not None:
            x = TF.adjust_saturation(x, p['sat'])
        if p['bri'] is not None:
            x = TF.adjust_brightness(x, p['bri'])
        if p['con'] is not None:
            x = TF.adjust_contrast(x, p['con'])
        if p['sigma'] is not None:
            x = TF.gaussian_blur(x, [5, 5], [p['sigma'], p['sigma']])
        if p['flip']:
            x = TF.hflip(x)
        d = (out[i] - x).abs()
        # hue: floor(h*6) may land on the other side of a sector boundary for a handful of pixels -> allow 1e-5 of them
        frac_bad = float((d > 2e-5).float().mean())
        assert frac_bad <= (1e-4 if p['hue'] is not None else 0.0), (i, frac_bad, float(d.max()))
    # sampling: frequencies of the decisions
    torch.manual_seed(0)
    s = [ClipAugment.sample() for _ in range(4000)]
    for k in ('hue', 'sat', 'bri', 'con', 'sigma'):
        assert 0.21 < sum(v[k] is not None for v in s) / 4000 < 0.29
    assert 0.45 < sum(v['flip'] for v in s) / 4000 < 0.55
    hs = [v['hue'] for v in s if v['hue'] is not None]
    assert -0.2 <= min(hs) and max(hs) <= 0.2 and 0.7 <= min(v['con'] for v in s if v['con'] is not None)
