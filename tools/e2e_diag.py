#!/usr/bin/env python3
"""Diagnostic: where the e2e arm of bench.py loses time against the device-resident arm (host pieces, crop DMA, steps)."""
import os
import sys
import time
import contextlib
import io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
import torch
import bench
from model.model import TDEEDModel
from tdeed_b200.pipeline import VideoInference, PendingEvents

dev = torch.device('cuda', 0)
with contextlib.redirect_stdout(io.StringIO()):
    model = TDEEDModel(device='cuda:0', args=bench.model_args())
bench.randomize_(model._model, 0)
model._model.eval()
eng = model._model.engine('bf16')
N = bench.VIDEO_FRAMES
starts = bench.clip_starts(N)
videos = [('video', N, starts)]
video = torch.randint(0, 256, (N, 3, bench.FRAME_H, bench.FRAME_W), dtype=torch.uint8, device=dev)
host = torch.empty(video.shape, dtype=torch.uint8).pin_memory()
host.copy_(video)
K = 5


def run(vi, src, piece, steps, post=True):
    pend = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    host_t = 0.0
    for _ in range(steps):
        h0 = time.perf_counter()
        vs = vi.run(videos, (src[lo:lo + piece] for lo in range(0, N, piece)))['video']
        ev = vs.events(0.01)
        if post:
            pend.append((PendingEvents(ev, K, 1, 0.01, False), PendingEvents(ev, K, 3, 0.01, True)))
        host_t += time.perf_counter() - h0
        while len(pend) > 1:
            a, b = pend.pop(0)
            a.get(), b.get()
    for a, b in pend:
        a.get(), b.get()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return len(starts) * steps / dt, host_t / steps * 1e3


for crop in (True,):
    vi = VideoInference(eng, (bench.FRAME_H, bench.FRAME_W), clips_per_batch=57, frames_per_chunk=1425, upload_crop=crop)
    run(vi, video, 1425, 2)
    for label, src, piece in (('device', video, 1425), ('host p=475', host, 475)):
        for steps in (20,):
            cps, ht = run(vi, src, piece, steps)
            print('crop=%s %-12s steps=%2d  %7.1f clips/s   host enqueue %.2f ms/video' % (crop, label, steps, cps, ht), flush=True)
    # raw H2D rate of one video through the same path, no compute
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for lo in range(0, N, 1425):
            vi._fill = 0
            vi._feed_only = True
            p = host[lo:lo + 1425]
            with torch.cuda.stream(vi.stream):
                if vi.upload_crop is None:
                    vi.bufs[0][:p.shape[0]].copy_(p, non_blocking=True)
                else:
                    from tdeed_b200.pipeline import _memcpy2d_async
                    _, x0, h, w = vi.upload_crop
                    _memcpy2d_async(vi.bufs[0].data_ptr(), w, p.data_ptr() + x0, bench.FRAME_W, w, p.shape[0] * 3 * bench.FRAME_H, vi.stream.cuda_stream)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    nbytes = N * 3 * bench.FRAME_H * (224 if crop else bench.FRAME_W)
    print('crop=%s raw H2D of one video: %.2f ms, %.1f GB/s' % (crop, dt * 1e3, nbytes / dt / 1e9), flush=True)

# ---- phase timing: events around every lower / upper graph launch, device vs host source
print('--- phase timing (ms, mean over 10 videos): time from the launch point of the previous phase to the end of this phase')
vi = VideoInference(eng, (bench.FRAME_H, bench.FRAME_W), clips_per_batch=57, frames_per_chunk=1425, upload_crop=True)
run(vi, video, 1425, 2)
run(vi, host, 475, 2)
orig_l, orig_u = eng.lower_graphed, eng.upper_graphed
marks = []


def wrap(tag, fn):
    def inner(*a, **k):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        marks.append((tag, e0, e1))
        return out
    return inner


eng.lower_graphed, eng.upper_graphed = wrap('lower', orig_l), wrap('upper', orig_u)
for label, src, piece in (('device', video, 1425), ('host', host, 475)):
    marks.clear()
    run(vi, src, piece, 10)
    torch.cuda.synchronize()
    tot = {}
    for tag, e0, e1 in marks:
        tot.setdefault(tag, []).append(e0.elapsed_time(e1))
    span = marks[0][1].elapsed_time(marks[-1][2]) / 10
    print(label, {k: round(sum(v) / len(v), 3) for k, v in tot.items()}, 'per-video span %.2f' % span,
          'sum of phases %.2f' % (sum(sum(v) for v in tot.values()) / 10))
eng.lower_graphed, eng.upper_graphed = orig_l, orig_u
