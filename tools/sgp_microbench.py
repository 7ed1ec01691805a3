#!/usr/bin/env python3
"""SGP temporal-layer microbench (BASELINE.json configs[4]): single SGPBlock, single SGPMixer and the
full ED-SGP-Mixer stack over B x T x C x (ks, r), against the HBM roofline for the bandwidth-bound token-mixing kernel and
against the tensor peak for the MLP / concat-fc GEMMs.

    python tools/sgp_microbench.py [--quick] [--out profiles/r1g_sgp_microbench.json]

Token mixing (`tdeed_sgp_mix_fwd`): algorithmic bytes = read x + write y + write the GEMM operand g
  = B*C*(t_in*4 + T*4 + T*sizeof(g)) (+ weights C*(2ks+up+8)*4).  MLP: 16*B*T*C^2 FLOPs; concat_fc: 12*B*T*C^2.
Timing: CUDA events around `reps` back-to-back launches after a warm-up; inputs are rotated over enough copies to exceed L2
when the tensors are small enough for that to matter (--flush).
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return p['hbm_gbs'], p['bf16_tflops_sustained']
    except Exception:
        return 6650.0, 1400.0


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--out', default=None)
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    from model.modules import EDSGPMIXERLayers, SGPBlock, SGPMixer
    from tdeed_b200 import ops
    from tdeed_b200.engine import sgp_up_size
    dev = torch.device('cuda')
    hbm, tf = peaks()
    Bs = [1, 4, 8, 32] if not args.quick else [8, 57]
    Ts = [50, 100, 200, 400, 800] if not args.quick else [100, 800]
    Cs = [368, 768]
    KR = [(ks, r) for ks in (3, 5, 7, 9, 11) for r in (2, 4)] if not args.quick else [(5, 4), (9, 4)]
    rows = []
    for C in Cs:
        for ks, r in KR:
            up = sgp_up_size(ks, r)
            torch.manual_seed(0)
            blk = SGPBlock(C, kernel_size=ks, k=r, init_conv_vars=0.1).to(dev).eval()
            mixer = SGPMixer(C, kernel_size=ks, k=r, init_conv_vars=0.1, t_size=100).to(dev).eval()
            for B in Bs:
                for T in Ts:
                    x = torch.randn((B, T, C), device=dev)
                    w = blk.mix_weights() if hasattr(blk, 'mix_weights') else None
                    rec = dict(C=C, ks=ks, r=r, up=up, B=B, T=T)
                    # (1) token mixing kernel alone (fp32 in, fp32 y + bf16 g out)
                    if w is not None:
                        t_mix = timed(lambda: ops.sgp_mix(x, T, ks, up, w, torch.bfloat16), args.reps)
                        nbytes = B * C * T * (4 + 4 + 2) + C * (2 * ks + up + 8) * 4
                        rec.update(mix_us=t_mix * 1e6, mix_gbs=nbytes / t_mix / 1e9, mix_frac_hbm=nbytes / t_mix / 1e9 / hbm)
                    # (2) whole SGPBlock (mix + MLP on tcgen05) in bf16
                    with torch.autocast('cuda', dtype=torch.bfloat16):
                        t_blk = timed(lambda: blk.forward_btc(x, T), args.reps)
                    flops = 16.0 * B * T * C * C
                    rec.update(block_us=t_blk * 1e6, block_mlp_tflops_if_all_gemm=flops / t_blk / 1e12)
                    # (3) SGPMixer token mixing alone: skip (T rows) + coarse (ceil(T/2) rows) -> 6C-wide bf16 concat
                    mw = mixer.mix_weights() if hasattr(mixer, 'mix_weights') else None
                    if mw is not None:
                        tc = (T + 1) // 2
                        xc = torch.randn((B, tc, C), device=dev)
                        t_mx = timed(lambda: ops.sgp_mixer_mix(xc, x, ks, up, mw, torch.bfloat16), args.reps)
                        nb = B * C * (T * 4 + tc * 4 + 6 * T * 2) + C * (4 * ks + 2 * up + 16) * 4
                        rec.update(mixer_us=t_mx * 1e6, mixer_gbs=nb / t_mx / 1e9, mixer_frac_hbm=nb / t_mx / 1e9 / hbm)
                    rec['launches'] = dict(mix=4, mixer=3)
                    rows.append(rec)
                    print(json.dumps(rec))
        # full encoder-decoder stack, reference config (n_layers 2, ks 9, r 4)
        for B in Bs:
            for T in Ts:
                torch.manual_seed(0)
                net = EDSGPMIXERLayers(C, T, num_layers=2, ks=9, k=4, concat=True).to(dev).eval()
                x = torch.randn((B, T, C), device=dev)
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    t_all = timed(lambda: net(x), max(3, args.reps // 4))
                rec = dict(stack='EDSGPMixer L=2 ks=9 r=4', C=C, B=B, T=T, us=t_all * 1e6, clips_per_s=B / t_all)
                rows.append(rec)
                print(json.dumps(rec))
    if args.out:
        json.dump(dict(hbm_gbs=hbm, bf16_tflops=tf, rows=rows), open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
