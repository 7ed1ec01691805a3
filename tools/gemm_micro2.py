#!/usr/bin/env python3
"""tdeed_gemm_fwd (tcgen05) on the stage-3/4 layer shapes of one 57-clip batch, with N variants (how much do the two n-tiles of
the N = 368 layers cost?).  L2 flushed between repetitions."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
import torch
from tdeed_b200 import _lib as L, ops

dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M4, M3 = 5700 * 49, 5700 * 196
SHAPES = [('s4.conv3 N=368', M4, 368, 368, True), ('s4.conv3 N=256', M4, 368, 256, True), ('s4.conv3 N=112', M4, 368, 112, True),
          ('s4.conv3 N=184', M4, 368, 184, True), ('s4.conv1 N=368', M4, 376, 368, False), ('s4.conv1 N=256', M4, 376, 256, False),
          ('s3.conv1', M3, 192, 152, False), ('s3.conv3', M3, 152, 152, True),
          ('K368 nores N368', M4, 368, 368, False), ('K368 nores N256', M4, 368, 256, False), ('K376 res N256', M4, 376, 256, True),
          ('K384 nores N256', M4, 384, 256, False), ('K320 nores N256', M4, 320, 256, False), ('K192 nores N256', M4, 192, 256, False)]
if os.environ.get('GEMM_SHAPE'):
    SHAPES = [SHAPES[int(os.environ['GEMM_SHAPE'])]]
for name, m, k, n, res in SHAPES:
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    b = torch.randn(n, device=dev)
    r = torch.randn(m, n, device=dev).to(torch.bfloat16) if res else None
    out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm([(a, k, 0, k)], w, b, residual=r, act=L.ACT_RELU, out=out, rows=m)
    for _ in range(50):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); us = ts[len(ts) // 2]
    byt = (m * k + m * n * (2 if res else 1) + n * k) * 2
    print('%-16s M=%8d K=%4d N=%4d  %8.1f us  %6.2f TB/s algorithmic' % (name, m, k, n, us, byt / us / 1e6), flush=True)
