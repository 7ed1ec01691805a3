#!/usr/bin/env python3
"""What makes the TMA-fed GEMM slow at K = 368: row pitch not a multiple of 128 B, the box origin, or the partial last k-block?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
import torch
from tdeed_b200 import _lib as L, ops

dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M4 = 5700 * 49
# name, lda, col0, K, N, ldo, residual(ldr)
CASES = [('K368 lda368', 368, 0, 368, 256, 256, 0), ('K368 lda384', 384, 0, 368, 256, 256, 0), ('K384 lda384', 384, 0, 384, 256, 256, 0),
         ('K384 lda400', 400, 0, 384, 256, 256, 0), ('K320 lda384 col0=64', 384, 64, 320, 256, 256, 0), ('K296 lda384 col0=88', 384, 88, 296, 256, 256, 0),
         ('K368 lda368 N368', 368, 0, 368, 368, 368, 0), ('K384 lda384 N384 ldo384', 384, 0, 384, 384, 384, 0),
         ('K384 lda384 N368 ldo384', 384, 0, 384, 368, 384, 0),
         ('K368 lda368 N368 res', 368, 0, 368, 368, 368, 368), ('K384 lda384 N384 res ld384', 384, 0, 384, 384, 384, 384)]
M3 = 5700 * 196
CASES3 = [('s3 K152 lda152 N152 res', 152, 0, 152, 152, 152, 152), ('s3 K128 lda152 N152 res', 152, 0, 128, 152, 152, 152),
          ('s3 K192 lda192 N152 res', 192, 0, 192, 152, 152, 152), ('s3 K152 lda152 N152', 152, 0, 152, 152, 152, 0),
          ('s3 K128 lda152 N152', 152, 0, 128, 152, 152, 0), ('s3 K192 lda192 N152', 192, 0, 192, 152, 152, 0),
          ('s3 K64 lda152 N152', 152, 0, 64, 152, 152, 0), ('s3 K160 lda160 N160', 160, 0, 160, 160, 160, 0),
          ('s3 K160 lda160 N160 res', 160, 0, 160, 160, 160, 160), ('s3 K152 N152 ldo160', 152, 0, 152, 152, 160, 0), ('s3 K152 N128', 152, 0, 152, 128, 128, 0), ('s3 K152 N64', 152, 0, 152, 64, 64, 0)]
if os.environ.get('S3'):
    CASES, M4 = CASES3, M3
for name, lda, col0, k, n, ldo, ldr in CASES:
    a = torch.randn(M4, lda, device=dev).to(torch.bfloat16)
    w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    b = torch.randn(n, device=dev)
    r = torch.randn(M4, ldr, device=dev).to(torch.bfloat16) if ldr else None
    out = torch.empty(M4, ldo, dtype=torch.bfloat16, device=dev)
    fn = lambda: ops.gemm([(a, lda, col0, k)], w, b, residual=(r[:, :n] if r is not None else None), act=L.ACT_RELU, out=out[:, :n], rows=M4)
    for _ in range(30):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print('%-30s %8.1f us' % (name, ts[len(ts) // 2]), flush=True)
