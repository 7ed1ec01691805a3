#!/usr/bin/env python3
"""Micro-benchmark of tdeed_gemm_fwd (tcgen05) on the backbone's layer shapes (dev tooling)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
import torch
from tdeed_b200 import _lib as L, ops

SHAPES = [  # name, M, K, N, residual
    ('s1.ds', 4076800, 32, 24, False), ('s1.conv3', 4076800, 24, 24, True), ('s2.conv1', 4076800, 24, 56, False),
    ('s2.conv3', 1019200, 56, 56, True), ('s3.b1.conv1', 1019200, 56, 152, False), ('s3.conv3', 254800, 152, 152, True),
    ('s4.b1.conv1', 254800, 152, 368, False), ('s4.conv3', 63700, 368, 368, True), ('sgp.mlp1', 1300, 368, 1472, False),
    ('sgp.mlp2', 1300, 1472, 368, True)]

def main():
    dev = torch.device('cuda')
    for name, m, k, n, res in SHAPES:
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        w = torch.randn(n, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, device=dev)
        r = torch.randn(m, n, device=dev).to(torch.bfloat16) if res else None
        out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            ops.gemm([(a, k, 0, k)], w, b, residual=r, act=L.ACT_RELU, out=out, rows=m)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 20
        for _ in range(reps):
            ops.gemm([(a, k, 0, k)], w, b, residual=r, act=L.ACT_RELU, out=out, rows=m)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        byt = (m * k + m * n * (2 if res else 1) + n * k) * 2
        print('%-12s M=%8d K=%4d N=%4d  %8.1f us  %6.2f TB/s  %7.1f TFLOP/s' % (name, m, k, n, us, byt / us / 1e6, 2.0 * m * n * k / us / 1e6))

if __name__ == '__main__':
    main()
