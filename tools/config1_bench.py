#!/usr/bin/env python3
"""BASELINE.json configs[0]: FineDiving_small (RegNetY-200MF + GSF or GSM, SGP enc-dec L=2 ks=7 r=4, K=5, displacement r=2,
clip_len 100, 224x224) forward on synthetic frames: the CPU oracle port on the host cores (B in {1, 4}, median of 5 after one
warm-up) next to the sm_100a engine (fp32 exact and bf16) on the same inputs, with the parity numbers.

    python tools/config1_bench.py [--out profiles/r2_config1.json]
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import sys
import time
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 't-deed_b200'), os.path.join(ROOT, 'oracle'), ROOT):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    import tdeed_oracle as O
    from model.model import TDEEDModel
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    recs = []
    for arch in ('rny002_gsf', 'rny002_gsm'):
        cfg = O.named_config('FineDiving_small', feature_arch=arch)
        sd = O.random_state(cfg, 0)
        margs = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=2, feature_arch=arch, clip_len=100,
                          n_layers=2, sgp_ks=7, sgp_r=4, num_classes=4, crop_dim=224)
        with contextlib.redirect_stdout(io.StringIO()):
            m = TDEEDModel(device='cuda', args=margs)
        m.load(sd)
        for B in (1, 4):
            torch.manual_seed(0)
            x = torch.randint(0, 256, (B, 100, 3, 224, 224), dtype=torch.uint8)
            if arch.endswith('gsm') and B == 4:
                cpu_s = None                              # the CPU leg is reported for B = 1 only for the GSM variant
            else:
                O.predict(sd, cfg, x)
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    _, ref = O.predict(sd, cfg, x)
                    ts.append(time.perf_counter() - t0)
                cpu_s = statistics.median(ts)
            rec = dict(config='FineDiving_small', feature_arch=arch, B=B, cpu_cores=cores,
                       cpu_oracle_s=cpu_s, cpu_oracle_clips_per_s=(B / cpu_s) if cpu_s else None)
            xd = x.cuda()
            for prec, amp in (('fp32', False), ('bf16', True)):
                m.predict(xd, use_amp=amp)
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    _, got = m.predict(xd, use_amp=amp)
                    ts.append(time.perf_counter() - t0)
                rec['gpu_%s_s' % prec] = statistics.median(ts)
                rec['gpu_%s_clips_per_s' % prec] = B / statistics.median(ts)
                if cpu_s:
                    rec['max_abs_prob_diff_%s_vs_cpu_oracle' % prec] = float(np.abs(got - ref).max())
            recs.append(rec)
            print(json.dumps(rec))
    if args.out:
        json.dump(recs, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
