"""GPU parity of every C-ABI kernel against the CPU oracle (oracle/tdeed_oracle.py) or, for the plain
GEMM, a torch fp32 matmul.  Tolerances: fp32 kernels 1e-4 relative-to-max (north star: 1e-3 end to end);
bf16 kernels are compared on bf16-rounded inputs with a tolerance that covers one bf16 output rounding."""
import math

import numpy as np
import pytest
import torch

import tdeed_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


GEMM_TOL_F32_OUT = 2e-5
GEMM_TOL_BF16_OUT = 6e-3


def _gemm_case(dev, dtype, m, n, ks, act, with_res, backend, gather=None, seed=0):
    from tdeed_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(seed)
    k = sum(ks)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dtype)
    bias = torch.randn(n, generator=g)
    segs_cpu, segs = [], []
    if gather:
        s, frames, h, wd = gather
        src = torch.randn(frames, h, wd, ks[0] + 8, generator=g).to(dtype)        # lda > k: extra columns ignored
        sub = src[:, ::s, ::s, :ks[0]].reshape(-1, ks[0])
        m = sub.shape[0]
        segs_cpu.append(sub)
        segs.append((src.to(dev), ks[0] + 8, 0, ks[0]))
    else:
        for i, kk in enumerate(ks):
            col0 = 8 * i                                                           # second segment starts mid-row
            lda = (col0 + kk + 7) // 8 * 8
            a = torch.randn(m, lda, generator=g).to(dtype)
            segs_cpu.append(a[:, col0:col0 + kk])
            segs.append((a.to(dev), lda, col0, kk))
    res = torch.randn(m, n, generator=g).to(dtype) if with_res else None
    ref = torch.cat([s_.float() for s_ in segs_cpu], 1) @ w.float().t() + bias
    if with_res:
        ref = ref + res.float()
    if act == L.ACT_RELU:
        ref = torch.relu(ref)
    elif act == L.ACT_GELU:
        ref = torch.nn.functional.gelu(ref)
    # the tcgen05 kernel stages residual and output in the same tile -> they share a dtype; a bf16 output adds one rounding
    out_dtype = dtype if (with_res and backend in (L.GEMM_TCGEN05, L.GEMM_TCGEN05_THIN)) else torch.float32
    out = ops.gemm(segs, w.to(dev), bias.to(dev), residual=res.to(dev) if with_res else None, act=act, rows=m,
                   out_dtype=out_dtype, backend=backend,
                   gather=(gather[0], gather[2], gather[3]) if gather else None)
    torch.cuda.synchronize()
    err = rel_err(out, ref)
    # explicit tolerances (relative to max|ref|): operands are bf16-exact, accumulation is fp32 -> an fp32 output must sit at
    # fp32 round-off; a bf16 output adds one round-to-nearest of the result (2^-9 relative) on top of the bf16 residual
    tol = GEMM_TOL_BF16_OUT if out_dtype == torch.bfloat16 else GEMM_TOL_F32_OUT
    assert err < tol, 'gemm m=%d n=%d ks=%s act=%d res=%s backend=%d: rel err %.3g >= %.3g' % (m, n, ks, act, with_res, backend, err, tol)
    return err


@pytest.mark.parametrize('m,n,ks', [(300, 24, (32,)), (1000, 152, (16, 40)), (70, 368, (92, 276)), (513, 1472, (368,))])
def test_gemm_simt_fp32(dev, m, n, ks):
    from tdeed_b200 import _lib as L
    for act, res in ((L.ACT_NONE, False), (L.ACT_RELU, True), (L.ACT_GELU, False)):
        _gemm_case(dev, torch.float32, m, n, ks, act, res, L.GEMM_SIMT)


def test_gemm_simt_gather(dev):
    from tdeed_b200 import _lib as L
    _gemm_case(dev, torch.float32, 0, 56, (24,), L.ACT_NONE, False, L.GEMM_SIMT, gather=(2, 3, 14, 10))
    _gemm_case(dev, torch.bfloat16, 0, 152, (56,), L.ACT_NONE, False, L.GEMM_SIMT, gather=(2, 2, 7, 9))


@pytest.mark.parametrize('m,n,ks', [
    (128, 32, (64,)), (300, 24, (32,)), (1000, 56, (24,)), (777, 152, (16, 40)), (4900, 368, (96, 280)),
    (400, 1472, (368,)), (400, 368, (1472,)), (1300, 768, (4608,)), (19600, 368, (368,)), (129, 3072, (768,))])
def test_gemm_tcgen05_bf16(dev, m, n, ks):
    """tcgen05/TMA kernel vs fp32 matmul of the same bf16 operands (fp32 accumulate -> only order differs)."""
    from tdeed_b200 import _lib as L
    for act, res in ((L.ACT_NONE, False), (L.ACT_RELU, True), (L.ACT_GELU, False)):
        _gemm_case(dev, torch.bfloat16, m, n, ks, act, res, L.GEMM_TCGEN05)


@pytest.mark.parametrize('m,n,ks', [(128, 32, (32,)), (300, 24, (24,)), (100000, 24, (32,)), (5000, 56, (24,)), (4097, 56, (56,)),
                                    (777, 152, (16, 40)), (1000, 256, (64,)), (333, 64, (40, 24)), (60000, 152, (16, 40))])
def test_gemm_tcgen05_thin_k(dev, m, n, ks):
    """Thin-K tcgen05 kernel (cp.async producers into the no-swizzle UMMA layout, resident weights)."""
    from tdeed_b200 import _lib as L
    for act, res in ((L.ACT_NONE, False), (L.ACT_RELU, True), (L.ACT_GELU, False)):
        _gemm_case(dev, torch.bfloat16, m, n, ks, act, res, L.GEMM_TCGEN05_THIN)


@pytest.mark.parametrize('n,k,gather', [(152, 56, (2, 3, 14, 10)), (56, 24, (2, 2, 7, 9)), (24, 32, (2, 1, 6, 300)),
                                         (368, 152, (2, 5, 14, 14)), (128, 64, (2, 3, 56, 100))])
def test_gemm_tcgen05_strided_gather(dev, n, k, gather):
    """Stride-2 1x1 shortcut conv as implicit GEMM: A is a 4D strided TMA view of the NHWC input."""
    from tdeed_b200 import _lib as L
    for act, res in ((L.ACT_NONE, False), (L.ACT_RELU, True)):
        _gemm_case(dev, torch.bfloat16, 0, n, (k,), act, res, L.GEMM_TCGEN05, gather=gather)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize('in_dtype', [torch.uint8, torch.float32])
@pytest.mark.parametrize('flip', [False, True])
def test_stem(dev, in_dtype, flip):
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(1)
    frames = torch.randint(0, 256, (5, 3, 50, 70), generator=g, dtype=torch.uint8)
    w = torch.randn(32, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(32, generator=g) * 0.1
    cfg = O.Config(crop_dim=None)
    crop = (3, 11, 44, 45)                                   # odd width -> ceil(w/2) outputs
    x = frames[None, :, :, 3:47, 11:56]
    xn = O.preprocess(x, cfg, flip=flip)
    ref = torch.relu(torch.nn.functional.conv2d(xn, w, b, stride=2, padding=1))
    out = ops.stem(frames.to(in_dtype).to(dev), crop, flip, w.to(dev), b.to(dev), torch.float32)
    assert rel_err(_nchw(out), ref) < 1e-5
    out16 = ops.stem(frames.to(in_dtype).to(dev), crop, flip, w.to(dev), b.to(dev), torch.bfloat16)
    assert rel_err(_nchw(out16), ref) < 6e-3


@pytest.mark.parametrize('in_dtype', [torch.uint8, torch.float32])
@pytest.mark.parametrize('flip,n1', [(False, 24), (True, 64), (False, 0)])
def test_stem_tcgen05_fused_conv1(dev, in_dtype, flip, n1):
    """tcgen05 stem (+ fused s1.b1.conv1) vs the fp32 oracle computed on the same bf16-rounded operands."""
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(8)
    frames = torch.randint(0, 256, (3, 3, 70, 90), generator=g, dtype=torch.uint8)
    w = (torch.randn(32, 3, 3, 3, generator=g) * 0.2).to(torch.bfloat16).float()
    b = torch.randn(32, generator=g) * 0.1
    crop = (2, 5, 67, 83)                                   # odd sizes: partial tiles in both directions
    x = frames[None, :, :, 2:69, 5:88]
    xn = O.preprocess(x, O.Config(crop_dim=None), flip=flip).to(torch.bfloat16).float()
    stem_ref = torch.relu(torch.nn.functional.conv2d(xn, w, b, stride=2, padding=1))
    w0 = torch.zeros(32, 32)
    w0[:, :27] = w.reshape(32, 27)
    args = (frames.to(in_dtype).to(dev), crop, flip, w0.to(torch.bfloat16).to(dev), b.to(dev))
    if n1 == 0:
        out, _ = ops.stem_tc(*args)
        assert rel_err(_nchw(out), stem_ref) < 6e-3
        return
    w1 = (torch.randn(n1, 32, generator=g) / math.sqrt(32)).to(torch.bfloat16).float()
    b1 = torch.randn(n1, generator=g) * 0.1
    c1_ref = torch.relu(torch.nn.functional.conv2d(stem_ref.to(torch.bfloat16).float(), w1[:, :, None, None], b1))
    w1p = torch.zeros((n1 + 15) // 16 * 16, 32)
    w1p[:n1] = w1
    sub, c1 = ops.stem_tc(*args, w1p.to(torch.bfloat16).to(dev), b1.to(dev), n1, True, 2)
    assert rel_err(_nchw(sub), stem_ref[:, :, ::2, ::2]) < 6e-3
    assert rel_err(_nchw(c1), c1_ref) < 8e-3


@pytest.mark.parametrize('flip,n1,geom', [(False, 24, (3, 70, 90, (2, 5, 67, 83))), (True, 64, (3, 70, 90, (2, 5, 67, 83))),
                                          (False, 24, (2, 224, 398, (0, 87, 224, 224))), (True, 24, (1, 96, 140, (0, 0, 96, 140)))])
def test_stem_raw_pixel_shifted_descriptor(dev, flip, n1, geom):
    """stem_tc2.cu (raw-pixel implicit GEMM, normalisation folded into the weights, tap pairs per MMA) vs the fp32 oracle."""
    from tdeed_b200 import ops
    nfr, H, W, crop = geom
    g = torch.Generator().manual_seed(8)
    frames = torch.randint(0, 256, (nfr, 3, H, W), generator=g, dtype=torch.uint8)
    w = torch.randn(32, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(32, generator=g) * 0.1
    cy, cx, h, wd = crop
    x = frames[None, :, :, cy:cy + h, cx:cx + wd]
    xn = O.preprocess(x, O.Config(crop_dim=None), flip=flip)
    stem_ref = torch.relu(torch.nn.functional.conv2d(xn, w, b, stride=2, padding=1))
    w1 = (torch.randn(n1, 32, generator=g) / math.sqrt(32)).to(torch.bfloat16).float()
    b1 = torch.randn(n1, generator=g) * 0.1
    c1_ref = torch.relu(torch.nn.functional.conv2d(stem_ref, w1[:, :, None, None], b1))
    w1p = torch.zeros((n1 + 15) // 16 * 16, 32)
    w1p[:n1] = w1
    wimg, b0, pad = ops.stem_tc2_weights(w.to(dev), b.to(dev))
    sub, c1 = ops.stem_tc2(frames.to(dev), crop, flip, wimg, b0, pad, w1p.to(torch.bfloat16).to(dev), b1.to(dev), n1, 2)
    assert rel_err(_nchw(sub), stem_ref[:, :, ::2, ::2]) < 1e-2
    assert rel_err(_nchw(c1), c1_ref) < 1.5e-2


@pytest.mark.parametrize('c,gw,stride,h,w', [(24, 8, 2, 20, 22), (152, 8, 1, 7, 9), (368, 8, 2, 14, 14),
                                              (64, 16, 2, 12, 10), (320, 16, 1, 5, 6)])
def test_conv3x3g(dev, c, gw, stride, h, w):
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, c, h, w, generator=g)
    wt = torch.randn(c, gw, 3, 3, generator=g) / math.sqrt(9 * gw)
    b = torch.randn(c, generator=g) * 0.1
    ref = torch.relu(torch.nn.functional.conv2d(x, wt, b, stride=stride, padding=1, groups=c // gw))
    out = ops.conv3x3g(_nhwc(x).to(dev), wt.to(dev), b.to(dev), gw, stride)
    assert rel_err(_nchw(out), ref) < 1e-5
    xb = x.to(torch.bfloat16)
    refb = torch.relu(torch.nn.functional.conv2d(xb.float(), wt, b, stride=stride, padding=1, groups=c // gw))
    outb = ops.conv3x3g(_nhwc(xb).to(dev), wt.to(dev), b.to(dev), gw, stride)
    assert rel_err(_nchw(outb), refb) < 6e-3


@pytest.mark.parametrize('c,gw,stride,h,w,n', [(24, 8, 2, 20, 22, 3), (24, 8, 2, 21, 19, 2), (56, 8, 2, 12, 12, 5),
                                                (152, 8, 1, 7, 9, 4), (368, 8, 2, 14, 14, 3), (368, 8, 1, 7, 7, 9),
                                                (64, 16, 2, 12, 10, 3), (320, 16, 1, 5, 6, 4), (768, 16, 1, 7, 7, 3)])
def test_conv3x3g_tcgen05(dev, c, gw, stride, h, w, n):
    """tcgen05 shifted-descriptor implicit GEMM vs F.conv2d on the same bf16-rounded operands."""
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(c, gw, 3, 3, generator=g) / math.sqrt(9 * gw)).to(torch.bfloat16).float()
    b = torch.randn(c, generator=g) * 0.1
    ref = torch.relu(torch.nn.functional.conv2d(x.float(), wt, b, stride=stride, padding=1, groups=c // gw))
    out = ops.conv3x3g_tc(_nhwc(x).to(dev), ops.conv3_weight_image(wt.to(dev), gw), b.to(dev), stride)
    assert tuple(out.shape) == (n, ref.shape[2], ref.shape[3], c)
    assert rel_err(_nchw(out), ref) < 6e-3


@pytest.mark.parametrize('c,rd,hw', [(24, 8, (9, 7)), (368, 92, (7, 7)), (768, 192, (4, 5))])
def test_se_and_pool(dev, c, rd, hw):
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, c, *hw, generator=g)
    w1, b1 = torch.randn(rd, c, generator=g) / math.sqrt(c), torch.randn(rd, generator=g) * 0.1
    w2, b2 = torch.randn(c, rd, generator=g) / math.sqrt(rd), torch.randn(c, generator=g) * 0.1
    s = x.mean((2, 3))
    s = torch.sigmoid(torch.relu(s @ w1.t() + b1) @ w2.t() + b2)
    ref = x * s[:, :, None, None]
    out = ops.se_(_nhwc(x).to(dev), w1.to(dev), b1.to(dev), w2.t().contiguous().to(dev), b2.to(dev))
    assert rel_err(_nchw(out), ref) < 1e-5
    te = torch.randn(2, c, generator=g)
    pooled = ops.pool_posenc(_nhwc(x).to(dev), 2, te.to(dev))
    assert rel_err(pooled, x.mean((2, 3)) + te.repeat(2, 1)) < 1e-5


@pytest.mark.parametrize('c,rd,hw,n', [(368, 92, (7, 7), 37), (160, 38, (14, 14), 16), (152, 14, (5, 5), 3), (768, 192, (4, 5), 20),
                                       (24, 8, (9, 7), 33), (56, 6, (6, 6), 17)])
def test_se_gate_bf16_tensor_core_fc(dev, c, rd, hw, n):
    """bf16 activations: the two SE 1x1 convolutions run on TF32 tensor-core tiles (11-bit significands, fp32 accumulate — the
    precision of the reference's fp16 autocast).  The gate must agree with the fp32 computation to ~1e-3, ragged frame counts
    (n % 16 != 0) and hidden widths that are not multiples of 8 included."""
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(c + rd)
    x = (torch.randn(n, c, *hw, generator=g) * 2).bfloat16().float()
    w1, b1 = torch.randn(rd, c, generator=g) / math.sqrt(c), torch.randn(rd, generator=g) * 0.1
    w2, b2 = torch.randn(c, rd, generator=g) / math.sqrt(rd), torch.randn(c, generator=g) * 0.1
    s = torch.sigmoid(torch.relu(x.mean((2, 3)) @ w1.t() + b1) @ w2.t() + b2)
    xh = _nhwc(x).to(dev).bfloat16()
    gate = ops.se_gate(xh, w1.to(dev), b1.to(dev), w2.t().contiguous().to(dev), b2.to(dev))
    assert gate.shape == (n, c) and float((gate.cpu() - s).abs().max()) < 2e-3
    out = ops.se_(xh.clone(), w1.to(dev), b1.to(dev), w2.t().contiguous().to(dev), b2.to(dev))
    assert torch.equal(out, (xh.float() * gate[:, None, None, :]).bfloat16())       # the scale pass applies exactly that gate


@pytest.mark.parametrize('c,rd,hw', [(368, 92, (7, 7)), (160, 38, (14, 14)), (24, 8, (6, 5))])
def test_se_gate_bits_do_not_depend_on_the_frame_count(dev, c, rd, hw):
    """A clip batch must equal its clips one by one: the gate of a frame may not depend on how many frames share the call (16-frame
    MMA tiles, kernel choices by frame count)."""
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(c)
    n = 16 * 148 + 23                                   # more than one wave of 16-frame CTAs, ragged last one
    x = torch.randn(n, hw[0], hw[1], c, generator=g).to(dev).bfloat16()
    w1, b1 = (torch.randn(rd, c, generator=g) / math.sqrt(c)).to(dev), (torch.randn(rd, generator=g) * 0.1).to(dev)
    w2t, b2 = (torch.randn(rd, c, generator=g) / math.sqrt(rd)).to(dev), (torch.randn(c, generator=g) * 0.1).to(dev)
    big = ops.se_gate(x, w1, b1, w2t, b2).clone()
    for lo, hi in ((0, 16), (100, 137), (n - 23, n)):
        small = ops.se_gate(x[lo:hi].contiguous(), w1, b1, w2t, b2)
        assert torch.equal(small, big[lo:hi])


@pytest.mark.parametrize('mode', ['gsf', 'gsm'])
@pytest.mark.parametrize('fold,c,hw,dtype', [(16, 56, (6, 5), 'f32'), (40, 152, (4, 4), 'f32'), (92, 368, (3, 2), 'f32'), (192, 768, (2, 3), 'f32'),
                                             # bf16 activations (the inference engine's path): the real layer shapes of rny002 / rny008
                                             (12, 56, (28, 50), 'bf16'), (36, 152, (14, 25), 'bf16'), (92, 368, (7, 13), 'bf16'),
                                             (32, 128, (9, 7), 'bf16'), (80, 320, (5, 6), 'bf16'), (192, 768, (2, 3), 'bf16'),
                                             (4, 16, (3, 3), 'bf16'),
                                             # wide frames (SoccerNetBall 448x796 -> 56x100 at stage 3): many row blocks per frame
                                             (32, 128, (6, 100), 'bf16')])
def test_gate_shift(dev, mode, fold, c, hw, dtype):
    from tdeed_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(4)
    clips, T = 2, 5
    x = torch.randn(clips * T, c, *hw, generator=g)
    if dtype == 'bf16':
        x = x.bfloat16().float()
    sd = {'g.bn.weight': torch.rand(fold, generator=g) + 0.5, 'g.bn.bias': torch.randn(fold, generator=g) * 0.1,
          'g.bn.running_mean': torch.randn(fold, generator=g) * 0.1, 'g.bn.running_var': torch.rand(fold, generator=g) + 0.5,
          'g.conv3D.weight': torch.randn(2, fold // 2, 3, 3, 3, generator=g) / math.sqrt(27 * fold / 2),
          'g.conv3D.bias': torch.randn(2, generator=g) * 0.1,
          'g.channel_conv1.weight': torch.randn(1, 2, 3, 3, generator=g) * 0.3, 'g.channel_conv1.bias': torch.randn(1, generator=g) * 0.1,
          'g.channel_conv2.weight': torch.randn(1, 2, 3, 3, generator=g) * 0.3, 'g.channel_conv2.bias': torch.randn(1, generator=g) * 0.1}
    ref = O.gate_shift(x[:, :fold], sd, 'g', T, mode)
    scale = sd['g.bn.weight'] / torch.sqrt(sd['g.bn.running_var'] + 1e-5)
    p = dict(bn_scale=scale, bn_shift=sd['g.bn.bias'] - sd['g.bn.running_mean'] * scale,
             w3d=sd['g.conv3D.weight'].reshape(-1), b3d=sd['g.conv3D.bias'],
             cc_w=torch.cat([sd['g.channel_conv1.weight'].reshape(-1), sd['g.channel_conv2.weight'].reshape(-1)]),
             cc_b=torch.cat([sd['g.channel_conv1.bias'], sd['g.channel_conv2.bias']]))
    p = {k: v.float().contiguous().to(dev) for k, v in p.items()}
    xh = _nhwc(x).to(dev)
    ld = (fold + 7) // 8 * 8
    ws = torch.empty(ops.gsf_workspace_floats(clips, T, hw[0], hw[1], fold), dtype=torch.float32, device=dev)
    if dtype == 'bf16':
        # pad columns are poisoned: the kernel must write them as zeros (the 1x1 conv reads them against zero weights)
        out = torch.full((clips * T * hw[0] * hw[1], ld), float('nan'), device=dev, dtype=torch.bfloat16)
        ops.gsf(xh.bfloat16(), clips, T, fold, L.SHIFT_GSF if mode == 'gsf' else L.SHIFT_GSM, p, ws, out)
        assert bool((out[:, fold:] == 0).all())
        got = out[:, :fold].float().reshape(clips * T, hw[0], hw[1], fold).permute(0, 3, 1, 2)
        # bf16 output rounding (2^-9) dominates; the gate conv runs on bf16 tensor-core operands with fp32 accumulation
        assert rel_err(got, ref) < 6e-3
        # natural channel order (the engine's call): the same values at the un-interleaved positions, bit for bit
        nat = torch.full_like(out, float('nan'))
        ops.gsf(xh.bfloat16(), clips, T, fold, L.SHIFT_GSF if mode == 'gsf' else L.SHIFT_GSM, p, ws, nat, natural=True)
        pos = ops.gsf_interleaved_positions(fold)
        assert sorted(pos) == list(range(fold))
        assert torch.equal(nat[:, :fold], out[:, pos]) and bool((nat[:, fold:] == 0).all())
        return
    out = torch.zeros(clips * T * hw[0] * hw[1], ld, device=dev)
    ops.gsf(xh, clips, T, fold, L.SHIFT_GSF if mode == 'gsf' else L.SHIFT_GSM, p, ws, out)
    got = out[:, :fold].reshape(clips * T, hw[0], hw[1], fold).permute(0, 3, 1, 2)
    assert rel_err(got, ref) < 2e-5
    nat = torch.zeros_like(out)
    ops.gsf(xh, clips, T, fold, L.SHIFT_GSF if mode == 'gsf' else L.SHIFT_GSM, p, ws, nat, natural=True)
    assert torch.equal(nat[:, :fold], out[:, ops.gsf_interleaved_positions(fold)])


@pytest.mark.parametrize('c,t_in,t_out,ks,r', [(368, 25, 25, 7, 4), (368, 25, 13, 5, 4), (768, 100, 50, 9, 4), (368, 13, 7, 11, 2),
                                              (768, 800, 800, 11, 4), (768, 800, 400, 3, 2)])   # last two: long-sequence (one-tile) mode
def test_sgp_block_fp32(dev, c, t_in, t_out, ks, r):
    from model.modules import SGPBlock
    torch.manual_seed(5)
    blk = SGPBlock(c, kernel_size=ks, k=r, init_conv_vars=0.1)
    for p in blk.parameters():
        p.data.add_(torch.randn_like(p) * 0.05)
    x = torch.randn(2, c, t_in)
    sd = {'b.' + k: v for k, v in blk.state_dict().items()}
    xp = torch.nn.functional.adaptive_max_pool1d(x, t_out) if t_out != t_in else x
    ref = O.sgp_block(xp, sd, 'b')
    blk = blk.to(dev).eval()
    got = blk.forward_btc(x.permute(0, 2, 1).contiguous().to(dev), t_out).permute(0, 2, 1)
    assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize('c,tc,t,ks,r', [(368, 13, 25, 7, 4), (768, 50, 100, 9, 4), (368, 7, 13, 5, 4), (768, 400, 800, 9, 4)])
def test_sgp_mixer_fp32(dev, c, tc, t, ks, r):
    from model.modules import SGPMixer
    torch.manual_seed(6)
    mix = SGPMixer(c, kernel_size=ks, k=r, init_conv_vars=0.1, t_size=t)
    for p in mix.parameters():
        p.data.add_(torch.randn_like(p) * 0.03)
    x, z = torch.randn(2, c, tc), torch.randn(2, c, t)
    sd = {'m.' + k: v for k, v in mix.state_dict().items()}
    ref = O.sgp_mixer(x, z, sd, 'm', t)
    mix = mix.to(dev).eval()
    got = mix(x.to(dev), z.to(dev))
    assert rel_err(got, ref) < 2e-5


def test_heads_softmax_scatter(dev):
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(7)
    for c, k_out, k_sm in ((368, 5, 5), (768, 31, 13), (768, 33, 33)):
        feat = torch.randn(3, 40, c, generator=g)
        w, b = torch.randn(k_out, c, generator=g) / math.sqrt(c) * 3, torch.randn(k_out, generator=g)
        wd, bd = torch.randn(c, generator=g) / math.sqrt(c) * 4, torch.randn(1, generator=g)
        logits_ref = feat @ w.t() + b
        displ_ref = feat @ wd + bd
        logits, displ, probs = ops.heads(feat.to(dev), w.to(dev), b.to(dev), wd.to(dev), bd.to(dev), k_sm)
        assert rel_err(logits, logits_ref) < 1e-5 and rel_err(displ, displ_ref) < 1e-5
        # scatter-max exactness: feed the kernel's own logits/displ to the oracle
        ref = O.scatter_max_probs(logits.cpu(), displ.cpu(), num_softmax=k_sm)
        assert float((probs.cpu() - ref).abs().max()) < 1e-6
        assert torch.equal(probs.cpu() == 0, ref == 0)          # untouched rows stay exactly zero
        again = ops.softmax_scatter(logits, displ, k_sm)
        assert torch.equal(again, probs)
        _, _, plain = ops.heads(feat.to(dev), w.to(dev), b.to(dev), None, None, k_sm)
        assert float((plain.cpu() - torch.softmax(logits.cpu()[..., :k_sm], 2)).abs().max()) < 1e-6
    # rounding: half-to-even and clamping
    lg = torch.zeros(1, 8, 3)
    lg[0, :, 1] = torch.arange(8.)
    d = torch.tensor([[0.5, 1.5, 2.5, -0.5, -1.5, 60.0, -60.0, 0.49]])
    got, ref = ops.softmax_scatter(lg.to(dev), d.to(dev), 3).cpu(), O.scatter_max_probs(lg, d)
    assert torch.equal(got == 0, ref == 0) and float((got - ref).abs().max()) < 1e-6


@pytest.mark.parametrize('frames,hw,k,n,with_res', [(40, 49, 368, 368, True), (23, 196, 152, 152, True), (9, 196, 152, 368, False),
                                                     (7, 30, 96, 88, True), (300, 49, 368, 368, True)])
def test_gemm_scaled_equals_scale_then_gemm(dev, frames, hw, k, n, with_res):
    """SE gate folded into conv3's A operand (tdeed_gemm_scaled_fwd) == stand-alone bf16(a * gate) followed by the plain
    tcgen05 GEMM, BIT FOR BIT (same fp32 multiply, same round-to-nearest bf16, same MMA order)."""
    from tdeed_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(frames + k)
    m = frames * hw
    a = torch.randn(m, k, generator=g).to(torch.bfloat16).to(dev)
    gate = torch.rand(frames, k, generator=g).to(dev)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(torch.bfloat16).to(dev)
    bias = torch.randn(n, generator=g).to(dev)
    res = torch.randn(m, n, generator=g).to(torch.bfloat16).to(dev) if with_res else None
    scaled = (a.float() * gate.repeat_interleave(hw, dim=0)).to(torch.bfloat16)
    want = ops.gemm([(scaled, k, 0, k)], w, bias, residual=res, act=L.ACT_RELU, rows=m, backend=L.GEMM_TCGEN05)
    got = ops.gemm_scaled(a, gate, hw, w, bias, residual=res, act=L.ACT_RELU)
    torch.cuda.synchronize()
    assert torch.equal(got, want), float((got.float() - want.float()).abs().max())
    ref = torch.relu(scaled.float() @ w.float().t() + bias + (res.float() if with_res else 0))
    assert rel_err(got.float(), ref) < GEMM_TOL_BF16_OUT


def test_engine_fused_se_gate_is_bit_identical_to_in_place_se(dev):
    from argparse import Namespace
    from model.model import TDEEDModel
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=8, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=4, radi_displacement=1, crop_dim=None)
    args = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=1, feature_arch=cfg.feature_arch, clip_len=8,
                     n_layers=2, sgp_ks=5, sgp_r=2, num_classes=4, crop_dim=None)
    m = TDEEDModel(device='cuda', args=args)
    m.load(O.random_state(cfg, 2))
    m._model.eval()
    eng = m._model.engine('bf16')
    frames = torch.randint(0, 256, (3, 8, 3, 64, 96), generator=torch.Generator().manual_seed(1), dtype=torch.uint8).to(dev)
    eng.fuse_se = True
    a = [t.clone() if t is not None else None for t in eng.forward(frames)]
    eng.fuse_se = False
    b = eng.forward(frames)
    for x, y in zip(a, b):
        assert (x is None and y is None) or torch.equal(x, y)
