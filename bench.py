#!/usr/bin/env python3
"""bench.py — T-DEED hot-path benchmark on B200 (contract: see DESIGN.md §Measurement).

Workload (BASELINE.json configs[1]): FigureSkatingComp_small (RegNetY-200MF + GSF, SGP enc-dec L=3 ks=5,
K=5, displacement head) batched inference + NMS.  One STEP = one synthetic video of 4 275 frames:
169 overlapping clips of 100 x 3 x 224 x 398 uint8 (center-cropped to 224^2 inside the stem kernel),
run in batches through the sm_100a engine, accumulated per video on the device, then event extraction,
NMS (window 1) and soft-NMS (window 3).  metric = clips/s (frames/s = 100 x).

  value : device-resident — the video's 4 275 unique frames already sit in HBM (1.14 GB > L2, so no L2 flush needed)
  e2e   : the same work through the host-facing path: pinned host uint8 frames -> H2D -> engine ->
          post-processing -> D2H of the event lists, all inside the timed region
The clips of a video overlap by 75 % (dataset/frame.py:409-423): the video-level engine (tdeed_b200.pipeline.VideoInference,
what util.eval.evaluate drives) runs stem + s1 + s2 once per unique frame and assembles the 169 clips from the cached
features — bit-identical to per-clip execution (tests/test_gpu_video.py).  `--path per-clip` measures the round-1 behaviour.
  --impl reference : the CPU oracle (port of the reference's PyTorch path) on the host cores

Multi-GPU (torchrun): clip-sharded inference, every rank processes its own videos (weak scaling, no
data-path collective); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIG = dict(name='FigureSkatingComp_small', feature_arch='rny002_gsf', clip_len=100, n_layers=3, sgp_ks=5, sgp_r=4,
              num_classes=4, radi_displacement=1, crop_dim=224)
FRAME_H, FRAME_W = 224, 398
VIDEO_FRAMES = 4275
NMS_WINDOW, SNMS_WINDOW = 1, 3


def clip_starts(num_frames, clip_len=100, overlap=75, stride=1, pad=5):
    return [s // stride for s in range(-pad * stride, max(0, num_frames - overlap * stride), (clip_len - overlap) * stride)]


def model_args():
    c = CONFIG
    return Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=c['radi_displacement'],
                     feature_arch=c['feature_arch'], clip_len=c['clip_len'], n_layers=c['n_layers'], sgp_ks=c['sgp_ks'],
                     sgp_r=c['sgp_r'], num_classes=c['num_classes'], crop_dim=c['crop_dim'])


def randomize_(module, seed=0):
    """Random-init weights of the right architecture with non-degenerate BN statistics / gammas (no pretrained
    checkpoint is available offline)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, buf in module.named_buffers():
            if name.endswith('running_mean'):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
            elif name.endswith('running_var'):
                buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
        for name, p in module.named_parameters():
            if '.bn.' in name and name.endswith('weight'):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif name.endswith('.bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        if not sm:
            return None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons, 'samples': len(sm)}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return dict(hbm=p['hbm_gbs'], tf=p['bf16_tflops_sustained'], src='measured')
    except Exception:
        return dict(hbm=6650.0, tf=1400.0, src='fallback')


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """CPU oracle (port of the reference's PyTorch fp32 path) on the host cores: 1 clip per step."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import tdeed_oracle as O
    import postproc_oracle as P
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = O.named_config(CONFIG['name'])
    sd = O.random_state(cfg, 0)
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (1, 100, 3, FRAME_H, FRAME_W), generator=g, dtype=torch.uint8)
    starts = clip_starts(VIDEO_FRAMES)
    scores = np.zeros((VIDEO_FRAMES, cfg.num_classes + 1), np.float32)
    support = np.zeros(VIDEO_FRAMES, np.int32)
    for _ in range(args.warmup):
        O.predict(sd, cfg, frames)
    t0 = time.perf_counter()
    for i in range(args.steps):
        _, probs = O.predict(sd, cfg, frames)
        P.accumulate_batched(scores, support, probs[0], starts[i % len(starts)])
    t_fwd = time.perf_counter() - t0
    # post-processing of one full video, charged pro rata (steps / clips-per-video)
    rng = np.random.default_rng(0)
    vs = rng.random((VIDEO_FRAMES, cfg.num_classes + 1)).astype(np.float32) ** 8
    t1 = time.perf_counter()
    _, _, hr = P.frame_predictions(vs, np.ones(VIDEO_FRAMES, np.int32), 0.01)
    P.nms(*hr, window=NMS_WINDOW, threshold=0.01)
    P.soft_nms(*hr, window=SNMS_WINDOW, threshold=0.01)
    t_post = (time.perf_counter() - t1) * args.steps / len(starts)
    total = t_fwd + t_post
    value = args.steps / total
    sample = '%d steps x 1 clip (100x3x224x398 u8) through the CPU oracle fp32 forward + numpy accumulate; ' \
             'NMS/SNMS of one 4275-frame video charged pro rata' % args.steps
    print(json.dumps({
        'impl': 'reference', 'metric': 'clips_per_s', 'value': value, 'unit': 'clips/s', 'frames_per_s': value * 100,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': CONFIG['name'] + ' batched inference + NMS', 'clips_per_step': 1},
        'cpu_baseline': {'value': value, 'unit': 'clips/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ------------------------------------------------------------------------------------------ our arm
class _SyntheticFrameReader:
    """dataset/frame.py:546-626 (FrameReaderVideo.load_frames) semantics over an in-memory synthetic video: used by
    `--path evaluate` so that the benchmark measures util.eval.evaluate itself, not a JPEG decoder."""

    def __init__(self, frames, lengths):
        self.frames, self.lengths = frames, lengths

    def load_frames(self, video_name, start, end, pad=False, stride=1, source_info=None):
        n = self.lengths[video_name]
        idx = [f for f in range(start, end, stride) if 0 <= f < n]
        n_pad_start = sum(1 for f in range(start, end, stride) if f < 0)
        n_pad_end = sum(1 for f in range(start, end, stride) if f >= n)
        if not idx:
            return -1
        ret = self.frames[idx[0]:idx[-1] + 1:stride]
        if n_pad_start > 0 or (pad and n_pad_end > 0):
            ret = torch.nn.functional.pad(ret, (0, 0, 0, 0, 0, 0, n_pad_start, n_pad_end if pad else 0))
        return ret


class SyntheticVideoDataset(torch.utils.data.Dataset):
    """The attribute surface of dataset/frame.py:385-517 (ActionSpotVideoDataset) that util.eval.evaluate touches."""

    def __init__(self, frames, names, clip_len=100, overlap_len=75, stride=1, pad_len=5, dataset='fs_comp', fps=25.0):
        n = frames.shape[0]
        self._labels = [{'video': v, 'num_frames': n, 'fps': fps,
                         'events': [{'frame': int(f), 'label': 'c%d' % (1 + i % 4)} for i, f in enumerate(range(37, n, 61))]}
                        for v in names]
        self._clip_len, self._stride, self._dataset = clip_len, stride, dataset
        self._frame_reader = _SyntheticFrameReader(frames, {v: n for v in names})
        self._clips = [(l['video'], i) for l in self._labels
                       for i in range(-pad_len * stride, max(0, n - overlap_len * stride), (clip_len - overlap_len) * stride)]

    def __len__(self):
        return len(self._clips)

    def __getitem__(self, idx):
        video, start = self._clips[idx]
        return {'video': video, 'start': start // self._stride,
                'frame': self._frame_reader.load_frames(video, start, start + self._clip_len * self._stride, pad=True, stride=self._stride)}

    def get_labels(self, video):
        meta = next(l for l in self._labels if l['video'] == video)
        out = np.zeros(meta['num_frames'], np.int64)
        for e in meta['events']:
            out[e['frame']] = int(e['label'][1:])
        return out

    @property
    def videos(self):
        return sorted((l['video'], l['num_frames'], l['fps']) for l in self._labels)

    @property
    def labels(self):
        return self._labels


def run_evaluate_path(args, model, dev, rank, world):
    """`--path evaluate`: what the reference's callers call (train_tdeed.py:193,263 -> util.eval.evaluate) on a synthetic
    dataset object, both flavours: batched (augment=False) and TTA (augment=True, the path of every non-SoccerNet set)."""
    import contextlib
    import io
    import util.eval as E
    g = torch.Generator().manual_seed(7)
    frames = torch.randint(0, 256, (VIDEO_FRAMES, 3, FRAME_H, FRAME_W), generator=g, dtype=torch.uint8)
    names = ['video%02d' % i for i in range(max(2, args.steps))]
    ds = SyntheticVideoDataset(frames, names)
    classes = {'c%d' % i: i for i in range(1, CONFIG['num_classes'] + 1)}
    out = {}
    for tag, augment in (('batched', False), ('tta', True)):
        for _ in range(2):              # warm-up: CUDA graphs of this chunk / batch geometry, the pinned host ring
            with contextlib.redirect_stdout(io.StringIO()):
                E.evaluate(model, ds, 'VAL', classes, printed=False, test=False, augment=augment)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            mAP = E.evaluate(model, ds, 'VAL', classes, printed=False, test=False, augment=augment)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[tag] = {'clips_per_s': len(ds) / dt, 'seconds': dt, 'videos': len(names), 'clips': len(ds), 'avg_mAP': float(mAP),
                    'views_per_clip': 2 if augment else 1}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--clips-per-batch', type=int, default=57)
    ap.add_argument('--frames-per-chunk', type=int, default=0, help='frames per lower() launch of the video-level engine (default: clips_per_batch * 25)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='infer', choices=['infer', 'train'],
                    help="infer: BASELINE configs[1] (headline); train: configs[2] FineGym_big training step (bench_train.py)")
    ap.add_argument('--path', default='engine', choices=['engine', 'evaluate', 'per-clip'],
                    help="engine: video-level engine (tdeed_b200.pipeline.VideoInference, what util.eval.evaluate drives); "
                         "evaluate: additionally time util.eval.evaluate on a synthetic dataset object; "
                         "per-clip: round-1 behaviour, every clip through the whole network (no frame-feature cache)")
    ap.add_argument('--no-aug', action='store_true', help='train workload: disable the per-clip torchvision augmentation')
    ap.add_argument('--ncu-step', action='store_true',
                    help='after the warm-up run ONE eager step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)')
    ap.add_argument('--no-train', action='store_true', help="infer workload: skip the short 'train_step' side measurement")
    ap.add_argument('--no-ref-gpu', action='store_true', help="skip the 'ref_gpu' leg (oracle port on stock torch, same GPU)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != 'reference' else args.warmup
    if args.workload == 'train':
        import bench_train
        if args.impl == 'reference':
            return bench_train.run_train_reference(args)
        out = bench_train.run_train(args)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
        return out
    if args.impl == 'reference':
        return run_reference(args)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1 and os.environ.get('TDEED_NUMA_BIND', '1') != '0':
        from tdeed_b200.parallel import bind_to_gpu_numa
        numa_cpus = bind_to_gpu_numa(local)          # before any pinned allocation: first touch places the pages
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from model.model import TDEEDModel
    from tdeed_b200 import ops
    from tdeed_b200.pipeline import PendingEvents, VideoInference, VideoScores
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)             # the conv / linear weights come from torch's default init: the same on every run and rank
        model = TDEEDModel(device='cuda:%d' % local, args=model_args())
    randomize_(model._model, seed=0)
    model._model.eval()
    eng = model._model.engine(args.precision)
    K = CONFIG['num_classes'] + 1
    B = args.clips_per_batch
    T, HOP = CONFIG['clip_len'], 25
    starts = clip_starts(VIDEO_FRAMES)
    n_clips = len(starts)                       # 169
    videos = [('video', VIDEO_FRAMES, starts)]
    per_clip = args.path == 'per-clip'

    # the synthetic video: its UNIQUE frames, resident in HBM (1.14 GB > L2) and in pinned host memory for the e2e arm
    video = torch.empty((VIDEO_FRAMES, 3, FRAME_H, FRAME_W), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    for lo in range(0, VIDEO_FRAMES, 475):
        video[lo:lo + 475] = torch.randint(0, 256, video[lo:lo + 475].shape, generator=gen, dtype=torch.uint8, device=dev)
    host_video = torch.empty(video.shape, dtype=torch.uint8).pin_memory()
    host_video.copy_(video)
    torch.cuda.synchronize()
    PIECE = 475                                  # host pieces of the e2e arm (3 per 1425-frame device chunk)

    if per_clip:
        vi = VideoInference(eng, (FRAME_H, FRAME_W), clips_per_batch=B, frames_per_chunk=B * T, flips=(False,))
    else:
        vi = VideoInference(eng, (FRAME_H, FRAME_W), clips_per_batch=B, frames_per_chunk=args.frames_per_chunk or B * HOP, flips=(False,))

    def pieces(src, n):
        for lo in range(0, VIDEO_FRAMES, n):
            yield src[lo:lo + n]

    def run_video(src, piece):
        """One video -> device-resident VideoScores.  Video-level engine: every unique frame through stem+s1+s2 once.
        per-clip (round-1 behaviour): every clip is materialised (zero padded) and pushed as 100 fresh frames."""
        if not per_clip:
            return vi.run(videos, pieces(src, piece))['video']
        def clip_frames():
            for s in starts:
                lo, hi = max(s, 0), min(s + T, VIDEO_FRAMES)
                c = src[lo:hi]
                if hi - lo < T:
                    pad = torch.zeros((T - (hi - lo),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
                    pad = pad.pin_memory() if not src.is_cuda else pad
                    c = torch.cat([pad, c] if s < 0 else [c, pad])
                    c = c.pin_memory() if not src.is_cuda else c
                yield c
        fake = [('clip%d' % i, T, [0]) for i in range(n_clips)]         # every clip its own 100-frame "video"
        per = vi.run(fake, clip_frames())
        vs = VideoScores(VIDEO_FRAMES, K, dev)
        for i, s in enumerate(starts):
            p = per['clip%d' % i].scores.view(1, T, K)
            vs.add(p, [s])
        return vs

    def step_device():
        vs = run_video(video, B * HOP)
        ev = vs.events(0.01)
        ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], K, NMS_WINDOW, 0.01, False)
        ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], K, SNMS_WINDOW, 0.01, True)

    d2h = [0]
    pending = []

    def step_e2e():
        """One video through the host-facing path.  Nothing here blocks the host: frame uploads run on a side stream
        (each unique frame crosses PCIe once), the event lists come back through async D2H copies that are collected one
        video later."""
        vs = run_video(host_video, PIECE)
        ev = vs.events(0.01)
        pending.append((PendingEvents(ev, K, NMS_WINDOW, 0.01, False), PendingEvents(ev, K, SNMS_WINDOW, 0.01, True)))
        while len(pending) > 1:                                           # D2H of the previous video's event lists
            a, b = pending.pop(0)
            d2h[0] = a.nbytes + b.nbytes
            a.get(), b.get()

    def drain_e2e():
        while pending:
            a, b = pending.pop(0)
            a.get(), b.get()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    POST_LAUNCHES = 1 + 2 * 3                    # extract_events + 2 x nms (3 kernels each)
    for _ in range(args.warmup):
        step_device()
    barrier()
    if args.ncu_step:                            # one eager step under cudaProfilerStart/Stop for ncu
        vi.use_graphs = False
        step_device()
        torch.cuda.synchronize()
        eng.prof = []
        torch.cuda.cudart().cudaProfilerStart()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        # launch-ordered family labels of the GEMM launches (one kernel launch per entry): tools/traffic_from_ncu.py joins them
        # with ncu's per-launch DRAM bytes
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        json.dump({'precision': args.precision, 'clips_per_batch': B, 'path': args.path, 'clips_per_step': n_clips,
                   'gemm_launches': [(lab, nb) for lab, _, nb, _, _ in eng.prof if lab in ('conv1x1', 'conv1x1_ds', 'sgp_gemm')],
                   'all_ops': [(lab, nb) for lab, _, nb, _, _ in eng.prof]},
                  open(os.path.join(ROOT, 'gpurun_out', 'profile_step_labels.json'), 'w'))
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches + vi.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    dev_s = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    launches = (eng.launches + vi.launches - l0) + args.steps * POST_LAUNCHES
    clocks = sampler.stop() if rank == 0 else None

    h2d0 = vi.h2d_bytes
    step_e2e()
    step_e2e()
    drain_e2e()
    h2d_per_step = (vi.h2d_bytes - h2d0) // 2
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e()
    drain_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    # per-kernel-family profile of one (eager) step with CUDA events on the launch stream
    roofline, families = None, {}
    if rank == 0:
        vi.use_graphs = False
        step_device()            # untimed: lets the caching allocator settle for the eager pass
        torch.cuda.synchronize()
        eng.prof = []
        torch.cuda._sleep(40_000_000)       # keep the GPU busy while the host enqueues, so event deltas are pure GPU time
        step_device()
        torch.cuda.synchronize()
        for label, flops, nbytes, a, b in eng.prof:
            f = families.setdefault(label, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            f['ms'] += a.elapsed_time(b)
            f['flops'] += flops
            f['bytes'] += nbytes
            f['launches'] += 1
        eng.prof = None
        vi.use_graphs = True
        pk = peaks()
        top = max(families, key=lambda k_: families[k_]['ms'])
        f = families[top]
        t_hbm, t_tc = f['bytes'] / (pk['hbm'] * 1e9), f['flops'] / (pk['tf'] * 1e12)
        sec = f['ms'] / 1e3
        if t_hbm >= t_tc:
            roofline = {'kernel': top, 'bound': 'hbm', 'achieved': f['bytes'] / sec / 1e9, 'peak': pk['hbm'], 'unit': 'GB/s'}
        else:
            roofline = {'kernel': top, 'bound': 'tensor', 'achieved': f['flops'] / sec / 1e12, 'peak': pk['tf'], 'unit': 'TFLOP/s'}
        roofline['frac'] = roofline['achieved'] / roofline['peak']
        roofline['traffic'] = None
        roofline['alg_bytes_per_launch'] = f['bytes'] / f['launches']
        try:    # measured DRAM bytes per launch of this family (ncu dram__bytes_read.sum + dram__bytes_write.sum, committed profile)
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
            if B == tr.get('clips_per_batch') and args.path == tr.get('path', 'engine') and top in tr['families']:
                roofline['traffic'] = tr['families'][top]['dram_bytes_per_launch']
                roofline['traffic_source'] = 'profiles/r2_traffic.json (ncu --set full, one eager step)'
        except Exception:
            pass
        roofline['peak_source'] = pk['src']
        roofline['ms_per_launch'] = f['ms'] / f['launches']
        roofline['share_of_step'] = f['ms'] / sum(v['ms'] for v in families.values())
        roofline['families'] = {k_: {'ms': round(v['ms'], 3), 'frac_hbm': round(v['bytes'] / (v['ms'] / 1e3) / 1e9 / pk['hbm'], 3),
                                     'frac_tensor': round(v['flops'] / (v['ms'] / 1e3) / 1e12 / pk['tf'], 3)}
                                for k_, v in sorted(families.items(), key=lambda kv: -kv[1]['ms']) if v['ms'] > 0}

    evaluate_path = None
    if args.path == 'evaluate' and world == 1:
        evaluate_path = run_evaluate_path(args, model, dev, rank, world)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import tdeed_oracle as O
        cores = os.cpu_count()
        prev_threads = torch.get_num_threads()
        torch.set_num_threads(cores)
        cfg = O.named_config(CONFIG['name'])
        sd = {k_: v.detach().cpu() for k_, v in model.state_dict().items()}
        x = video[:100].cpu().unsqueeze(0)
        O.predict(sd, cfg, x)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            O.predict(sd, cfg, x)
        dt = (time.perf_counter() - t0) / reps
        cpu_baseline = {'value': 1.0 / dt, 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
                        'sample': '%d x 1 clip (100x3x224x398 u8) through the CPU oracle fp32 forward, torch threads = %d' % (reps, cores)}
        # the training side measurement below enqueues from this thread: do not leave a wide intra-op pool spinning next to it
        torch.set_num_threads(max(1, min(prev_threads, 2)))

    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_ref_gpu:
        del video
        vi = None
        eng._graphs.clear()
        torch.cuda.empty_cache()
        try:
            import bench_ref_gpu
            sd = {k_: v.detach() for k_, v in model.state_dict().items()}
            ref_gpu = {'inference': bench_ref_gpu.run_inference(CONFIG['name'], sd, (FRAME_H, FRAME_W), dev)}
            torch.cuda.empty_cache()
        except Exception as exc:
            ref_gpu = {'error': repr(exc)[:300]}

    if rank == 0:
        total_clips = n_clips * args.steps * world
        value = total_clips / dev_s
        out = {
            'metric': 'clips_per_s', 'value': value, 'unit': 'clips/s', 'frames_per_s': value * 100, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_s / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': CONFIG['name'] + ' batched inference + NMS', 'clips_per_step': n_clips,
                       'clips_per_batch': B, 'path': args.path, 'frame_shape': [100, 3, FRAME_H, FRAME_W], 'video_frames': VIDEO_FRAMES,
                       'unique_frames_per_step': VIDEO_FRAMES if not per_clip else n_clips * T,
                       'l2_policy': 'inputs (1.14 GB of unique frames per video) larger than L2', 'parallelism': 'clip-sharded x%d' % world,
                       'numa_bound_cpus': (len(numa_cpus) if numa_cpus else None)},
            'e2e': {'value': total_clips / e2e_s, 'unit': 'clips/s', 'h2d_bytes_per_step': int(h2d_per_step),
                    'd2h_bytes_per_step': int(d2h[0])},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
            'kernel_families_ms_per_step': {k_: round(v['ms'], 3) for k_, v in sorted(families.items(), key=lambda kv: -kv[1]['ms'])},
            'ref_gpu': ref_gpu, 'evaluate_path': evaluate_path,
        }
    # side measurement: the training step (BASELINE configs[2]) at this N, a few steps — reported under 'train_step'
    train = None
    if not args.no_train:
        video = host_video = vi = None
        eng._graphs.clear()
        model._model.__dict__.pop('_video_inference', None)
        torch.cuda.empty_cache()
        try:
            import bench_train
            targs = Namespace(steps=10, warmup=3, precision=args.precision, no_aug=False, gpus=args.gpus, no_ref_gpu=args.no_ref_gpu)
            tr = bench_train.run_train(targs, quiet=True)
            if tr is not None:
                train = {k_: tr.get(k_) for k_ in ('metric', 'value', 'unit', 'ms_per_step', 'config', 'e2e', 'gpu_launches', 'roofline',
                                                   'kernel_families_ms_per_step', 'ref_gpu')}
        except Exception as exc:      # the headline line must survive a failure of the side measurement
            train = {'error': repr(exc)[:300]}
    if rank == 0:
        out['train_step'] = train
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
