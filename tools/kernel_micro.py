#!/usr/bin/env python3
"""Stand-alone timings (CUDA events, L2 flushed between repetitions) of the small per-frame kernels at the shapes of one
57-clip batch of the bench workload: SE (mean / fc / scale), gate-shift, pool.  usage: python tools/kernel_micro.py [se gsf]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
import torch
from tdeed_b200 import ops, _lib as L

dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    import time
    t0 = time.time()
    while time.time() - t0 < 0.5:       # warm the clocks
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


what = sys.argv[1:] or ['se', 'gsf']
N = 5700
if 'se' in what:
    for (hw, c, rd, n) in ((49, 368, 92, N), (196, 152, 38, N), (196, 152, 14, N), (49, 368, 38, N), (784, 56, 6, 1425), (3136, 24, 8, 1425)):
        x = torch.randn(n, hw, 1, c, device=dev).to(torch.bfloat16)
        w1 = torch.randn(rd, c, device=dev) * 0.1
        b1 = torch.randn(rd, device=dev) * 0.1
        w2 = torch.randn(rd, c, device=dev) * 0.1
        b2 = torch.randn(c, device=dev) * 0.1
        tg = timeit(lambda: ops.se_gate(x, w1, b1, w2, b2))
        print('se_gate n=%d hw=%d c=%d rd=%d: %.1f us (mean+fc)' % (n, hw, c, rd, tg), flush=True)
        t = timeit(lambda: ops.se_(x, w1, b1, w2, b2))
        print('se   n=%d hw=%d c=%d rd=%d: %.1f us total (mean+fc+scale), %.2f TB/s on 3 passes' % (n, hw, c, rd, t, 3 * x.numel() * 2 / t / 1e6), flush=True)
if 'gsf' in what:
    shapes = ((28, 56, 16), (14, 152, 40), (7, 368, 92))
    if os.environ.get('GSF_SHAPE'):
        shapes = (shapes[int(os.environ['GSF_SHAPE'])],)
    for (h, c, fold) in shapes:
        b, t_ = 57, 100
        x = torch.randn(b * t_, h, h, c, device=dev).to(torch.bfloat16)
        p = dict(bn_scale=torch.rand(fold, device=dev) + 0.5, bn_shift=torch.randn(fold, device=dev) * 0.1,
                 w3d=(torch.randn(2 * (fold // 2) * 27, device=dev) * 0.05), b3d=torch.randn(2, device=dev) * 0.1,
                 cc_w=torch.randn(36, device=dev) * 0.2, cc_b=torch.randn(2, device=dev) * 0.1)
        ws = torch.empty(ops.gsf_workspace_floats(b, t_, h, h, fold), dtype=torch.float32, device=dev)
        out = torch.empty((b * t_ * h * h, (fold + 7) // 8 * 8), dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.gsf(x, b, t_, fold, L.SHIFT_GSF, p, ws, out, natural=True))
        print('gsf  %dx%dx%d fold=%d: %.1f us (q+gate+weight+blend)' % (h, h, c, fold, t), flush=True)
if 'c3' in what:
    # grouped 3x3 conv (tcgen05) at the layer shapes of one 57-clip batch (upper) and one 1425-frame chunk (lower)
    shapes = ((N, 14, 160, 8, 1), (N, 7, 368, 8, 1), (N, 28, 152, 8, 2), (N, 14, 368, 8, 2), (1425, 56, 56, 8, 2), (1425, 112, 24, 8, 2))
    if os.environ.get('C3_FAMILY') == 'rny008':
        shapes = ((N, 14, 320, 16, 1), (N, 7, 768, 16, 1), (N, 28, 320, 16, 2), (N, 14, 768, 16, 2), (1425, 28, 128, 16, 1),
                  (1425, 56, 128, 16, 2), (1425, 112, 64, 16, 2))
    if os.environ.get('C3_SHAPE'):
        shapes = (shapes[int(os.environ['C3_SHAPE'])],)
    for (n, h, c, gw, stride) in shapes:
        x = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16)
        wimg = ops.conv3_weight_image(torch.randn(c, gw, 3, 3, device=dev) * 0.1, gw)
        bias = torch.randn(c, device=dev) * 0.1
        ho = (h + stride - 1) // stride
        out = torch.empty((n, ho, ho, c), dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.conv3x3g_tc(x, wimg, bias, stride, out))
        nbytes = x.numel() * 2 + out.numel() * 2
        print('c3   n=%d %dx%dx%d s%d: %.1f us, %.2f TB/s algorithmic' % (n, h, h, c, stride, t, nbytes / t / 1e6), flush=True)
