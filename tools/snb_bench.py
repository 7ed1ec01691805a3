#!/usr/bin/env python3
"""BASELINE.json configs[3]: SoccerNetBall challenge config (RegNetY-800MF + GSF, double head 13+18, displacement head,
uncropped 448x796 frames, stride 2) — clip-sharded inference over a synthetic match.  One rank per GPU (torchrun) takes a
contiguous share of the 2 861 clips of a 143 188-frame match; scores are accumulated on the owning GPU, NMS (window 6) and
soft-NMS (window 12) run on the device.  `--clips N` bounds the number of clips per rank (default 64) so the run stays short;
the printed clips/s is per-job (sum over ranks).

    python tools/snb_bench.py [--clips 64] [--batch 4] [--arch rny008_gsf]
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--clips', type=int, default=64)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--arch', default='rny008_gsf')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from bench import randomize_
    from model.model import TDEEDModel
    from tdeed_b200 import ops
    from tdeed_b200.pipeline import VideoScores
    margs = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=4, feature_arch=args.arch, clip_len=100,
                      n_layers=2, sgp_ks=9, sgp_r=4, num_classes=12, crop_dim=-1)
    with contextlib.redirect_stdout(io.StringIO()):
        model = TDEEDModel(device='cuda:%d' % local, args=margs)
    model._model.update_pred_head([13, 18])
    model._num_classes = 31
    randomize_(model._model, seed=0)
    model._model.eval()
    eng = model._model.engine('bf16')
    H, W, stride = 448, 796, 2
    num_frames = 143188
    video_len = num_frames // stride
    starts_all = [s // stride for s in range(-5 * stride, max(0, num_frames - 50 * stride), (100 - 50) * stride)]     # 2 861 clips
    per = (len(starts_all) + world - 1) // world
    mine = starts_all[rank * per:(rank + 1) * per][:args.clips]
    B = args.batch
    gen = torch.Generator(device=dev).manual_seed(rank)
    clips = torch.randint(0, 256, (B, 100, 3, H, W), generator=gen, dtype=torch.uint8, device=dev)      # 428 MB, reused for every batch
    K = 13

    def run():
        vs = VideoScores(video_len, K, dev)
        for i in range(0, len(mine), B):
            st = mine[i:i + B]
            _, _, probs = eng.forward_graphed(clips[:len(st)])
            vs.add(probs, st)
        ev = vs.events(0.01)
        ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], K, 6, 0.01, False)
        ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], K, 12, 0.01, True)

    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    if rank == 0:
        total = len(mine) * world
        print(json.dumps({'workload': 'SoccerNetBall challenge (%s, 448x796, double head) clip-sharded inference' % args.arch,
                          'n_gpus': world, 'clips': total, 'clips_per_batch': B, 'seconds': sec, 'clips_per_s': total / sec,
                          'frames_per_s': total * 100 / sec, 'tflops_per_s': total / sec * (1147.7e9 if args.arch.startswith('rny008') else 287.7e9) / 1e12,
                          'full_match_seconds_at_this_rate': len(starts_all) / (total / sec)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
