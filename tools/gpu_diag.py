#!/usr/bin/env python3
"""GPU diagnostic: per-layer error of the sm_100a engine against the CPU oracle (test infrastructure).

    python tools/gpu_diag.py [case] > gpurun_out/diag.txt

Prints, for fp32 and bf16 engines, the relative-to-max error of every tapped activation (stem, each
bottleneck, pooled features, temporal output, logits, displacement) so a wrong kernel can be located
from a single gpurun call.
"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 't-deed_b200'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)

import torch  # noqa: E402

import tdeed_oracle as O  # noqa: E402
from tdeed_b200.engine import EngineConfig, InferenceEngine  # noqa: E402
from tdeed_b200 import _lib as L  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def main():
    kw = dict(feature_arch=sys.argv[1] if len(sys.argv) > 1 else 'rny002_gsf', clip_len=12, n_layers=2, sgp_ks=7,
              sgp_r=4, num_classes=4, radi_displacement=2, crop_dim=64)
    cfg = O.Config(**kw)
    sd = O.random_state(cfg, 1)
    g = torch.Generator().manual_seed(2)
    frames = torch.randint(0, 256, (2, 12, 3, 64, 80), generator=g, dtype=torch.uint8)
    taps_ref = {}
    with torch.no_grad():
        logits_ref, displ_ref = O.forward(sd, cfg, frames, taps=taps_ref)
    ecfg = EngineConfig(cfg.feature_arch, cfg.clip_len, cfg.n_layers, cfg.sgp_ks, cfg.sgp_r, cfg.num_classes,
                        cfg.radi_displacement, cfg.crop_dim)
    for precision, backend in (('fp32', L.GEMM_AUTO), ('bf16', L.GEMM_SIMT), ('bf16', L.GEMM_AUTO)):
        print('==== precision', precision, 'gemm backend', backend, flush=True)
        try:
            eng = InferenceEngine(ecfg, sd, precision=precision, gemm_backend=backend)
            taps = {}
            t0 = time.time()
            logits, displ, probs = eng.forward(frames.cuda(), taps=taps)
            torch.cuda.synchronize()
            print('forward ok in %.3fs, launches %d' % (time.time() - t0, eng.launches))
            for k, v in taps.items():
                r = taps_ref.get(k if k != 'feat_posenc' else 'feat')
                if r is None:
                    continue
                if k == 'feat_posenc':
                    r = r + sd['temp_enc'][None]
                if v.dim() == 4:
                    v = v.permute(0, 3, 1, 2)
                print('  %-14s rel_err %.3e   ref absmax %.3f' % (k, rel(v.reshape(r.shape), r), float(r.abs().max())))
            print('  %-14s rel_err %.3e' % ('logits', rel(logits, logits_ref)))
            print('  %-14s rel_err %.3e' % ('displ', rel(displ, displ_ref)))
        except Exception:
            traceback.print_exc()
        sys.stdout.flush()


if __name__ == '__main__':
    main()
