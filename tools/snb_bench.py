#!/usr/bin/env python3
"""BASELINE.json configs[3]: SoccerNetBall challenge config (RegNetY-800MF + GSF — challenge2 — or RegNetY-200MF — challenge1;
double head 13+18, displacement head radius 4, uncropped 448x796 frames, stride 2) — clip-sharded inference over a synthetic
match of 143 188 frames (data/soccernetball/challenge.json): video_len 71 594, clip starts range(-10, 143 038, 50) = 2 861
clips of 100 frames (dataset/frame.py:409-417 with overlap 75, stride 2).  One rank per GPU (torchrun) takes a contiguous share
of the match; scores are accumulated on the owning GPU, NMS (window 6) and soft-NMS (window 12) run on the device.

    python tools/snb_bench.py [--arch rny008_gsf|rny002_gsf] [--path engine|per-clip] [--frames N] [--batch B] [--out file.json]

--path engine  : the video-level engine (every unique frame through stem + s1 + s2 once; what util.eval.evaluate drives)
--path per-clip: every clip through the whole network (the reference's schedule)
--frames N     : only the first N (stride-2) frames of the match (default: the whole match); the rate is per job
Frames come from a 400-frame pinned pool that is cycled (the match itself would be 76 GB of uint8): host -> device copies,
post-processing and the D2H of the event lists are inside the timed region.
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
    ap.add_argument('--arch', default='rny008_gsf')
    ap.add_argument('--path', default='engine', choices=['engine', 'per-clip'])
    ap.add_argument('--frames', type=int, default=0)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--out', default=None)
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
    from tdeed_b200.pipeline import PendingEvents, VideoInference
    margs = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=4, feature_arch=args.arch, clip_len=100,
                      n_layers=2, sgp_ks=9, sgp_r=4, num_classes=12, crop_dim=-1)
    with contextlib.redirect_stdout(io.StringIO()):
        model = TDEEDModel(device='cuda:%d' % local, args=margs)
    model._model.update_pred_head([13, 18])
    model._num_classes = 31
    randomize_(model._model, seed=0)
    model._model.eval()
    eng = model._model.engine('bf16')
    H, W, stride, T, hop, K = 448, 796, 2, 100, 25, 13
    num_frames = 143188
    match_len = -(-num_frames // stride)                                  # 71 594
    starts_all = [s // stride for s in range(-5 * stride, max(0, num_frames - 75 * stride), (T - 75) * stride)]     # 2 861 clips
    assert len(starts_all) == 2861, len(starts_all)
    if args.frames:
        match_len = min(match_len, args.frames)
        starts_all = [s for s in starts_all if s < match_len - 75]
    # contiguous share of the match per rank: frames [lo, hi) and the clips that start in it (boundary clips read past hi: they
    # are padded here; a production run hands the <= 3 boundary clips' frames to both neighbours)
    per = -(-len(starts_all) // world)
    mine = starts_all[rank * per:(rank + 1) * per]
    f_lo = max(0, mine[0])
    f_hi = min(match_len, mine[-1] + T)
    local_starts = [s - f_lo for s in mine]
    vlen = f_hi - f_lo
    B = args.batch
    pool = torch.randint(0, 256, (400, 3, H, W), generator=torch.Generator().manual_seed(rank), dtype=torch.uint8).pin_memory()

    from tdeed_b200.pipeline import VideoScores
    engine_path = args.path == 'engine'
    vi = VideoInference(eng, (H, W), clips_per_batch=B, frames_per_chunk=B * (hop if engine_path else T), flips=(False,))

    def run(n_frames, starts):
        """frames [0, n_frames) of this rank's share and the clips `starts` (relative) -> (#events after NMS, after SNMS)"""
        if engine_path:
            def pieces():
                done = 0
                while done < n_frames:
                    n = min(100, n_frames - done)
                    o = (done // 100 * 100) % 400
                    yield pool[o:o + n]
                    done += n
            vs = vi.run([('match', n_frames, starts)], pieces())['match']
        else:
            per = vi.run([('c%d' % i, T, [0]) for i in range(len(starts))],
                         (pool[(i * 25) % 300:(i * 25) % 300 + T] for i in range(len(starts))))
            vs = VideoScores(n_frames, K, dev)
            for i, s in enumerate(starts):
                vs.add(per['c%d' % i].scores.view(1, T, K), [s])
        ev = vs.events(0.01)
        a, b = PendingEvents(ev, K, 6, 0.01, False), PendingEvents(ev, K, 12, 0.01, True)
        return len(a.get()[0]), len(b.get()[0])

    warm_frames = min(vlen, 3 * B * hop)
    run(warm_frames, [s for s in local_starts if s < warm_frames - 75][:3 * B])      # captures the CUDA graphs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_nms, n_snms = run(vlen, local_starts)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    if rank == 0:
        total = len(starts_all)
        gflop = 1147.7 if args.arch.startswith('rny008') else 287.7
        rec = {'workload': 'SoccerNetBall challenge (%s, 448x796 uint8, double head 13+18, displacement r=4), clip-sharded inference + NMS/SNMS' % args.arch,
               'path': args.path, 'n_gpus': world, 'clips': total, 'frames': match_len, 'clips_per_batch': B, 'seconds': sec,
               'clips_per_s': total / sec, 'frames_per_s': total * 100 / sec, 'nominal_tflops_per_s': total / sec * gflop / 1e3,
               'full_match_seconds_at_this_rate': 2861 / (total / sec), 'events_after_nms': n_nms, 'events_after_snms': n_snms,
               'h2d_bytes': int(vi.h2d_bytes), 'timing': 'wall clock incl. H2D of the frames and D2H of the event lists'}
        print(json.dumps(rec))
        if args.out:
            prev = []
            if os.path.exists(args.out):
                prev = json.load(open(args.out))
            json.dump(prev + [rec], open(args.out, 'w'), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
