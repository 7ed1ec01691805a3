"""Device-resident video inference: clips -> engine forward -> per-video score accumulation -> event
extraction -> NMS / soft-NMS, without host round trips between the stages.

This is the B200-native shape of util/eval.py:264-419 of the reference (`evaluate`): the reference
copies every batch of probabilities to the host and accumulates / suppresses in numpy and pure Python;
here per-video `scores` / `support` stay in HBM and only the final event lists are read back.
"""
import numpy as np
import torch

from . import ops


class VideoScores:
    """scores float32 (video_len, K) and support int32 (video_len) of one video, on the device."""

    def __init__(self, video_len, num_classes_p1, device):
        self.video_len = video_len
        self.k = num_classes_p1
        self.scores = torch.zeros((video_len, num_classes_p1), dtype=torch.float32, device=device)
        self.support = torch.zeros(video_len, dtype=torch.int32, device=device)

    def add(self, probs, starts, tta=False):
        """probs (n_clips, T, K) fp32 device tensor; starts: list/array of clip start frames (already // stride).
        Clip order == the reference's accumulation order (bit-exact fp32 sums)."""
        ops.clip_accumulate(self.scores, self.support, probs.contiguous(), [int(s) for s in starts], 1 if tta else 0)

    def events(self, threshold=0.01):
        """Normalise in place and extract events (util/eval.py:87-193).  Returns the device-side dict of
        tdeed_b200.ops.extract_events."""
        return ops.extract_events(self.scores, self.support, threshold)


_cudart = None


def _memcpy2d_async(dst_ptr, dpitch, src_ptr, spitch, width, height, stream):
    """cudaMemcpy2DAsync(host -> device) through the CUDA runtime PyTorch already loaded (plumbing, not compute)."""
    global _cudart
    if _cudart is None:
        import ctypes
        import glob
        import os
        lib = None
        for soname in ('libcudart.so.12', 'libcudart.so.13', 'libcudart.so'):   # the copy PyTorch already mapped
            try:
                lib = ctypes.CDLL(soname)
                break
            except OSError:
                pass
        for pat in () if lib is not None else (os.path.join(os.path.dirname(torch.__file__), 'lib', 'libcudart*.so*'),
                    os.path.join(os.path.dirname(torch.__file__), '..', 'nvidia', 'cuda_runtime', 'lib', 'libcudart.so*'),
                    '/usr/local/cuda/lib64/libcudart.so*'):
            hits = sorted(glob.glob(pat))
            if hits:
                lib = ctypes.CDLL(hits[0])
                break
        if lib is None:
            raise RuntimeError('libcudart not found')
        lib.cudaMemcpy2DAsync.restype = ctypes.c_int
        lib.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                          ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        _cudart = lib
    rc = _cudart.cudaMemcpy2DAsync(dst_ptr, dpitch, src_ptr, spitch, width, height, 1, stream)   # 1 = cudaMemcpyHostToDevice
    if rc != 0:
        raise RuntimeError('cudaMemcpy2DAsync failed with cudaError %d' % rc)


class ClipUploader:
    """Double-buffered host->device staging of clip batches on a side stream, so the PCIe copy of batch i+1
    overlaps the kernels of batch i (the reference blocks on `.to(device)` inside predict, model/model.py:340-342).

        up = ClipUploader(batch_shape, device)
        for host_clips in batches:                 # pinned uint8 (B,T,3,H,W) tensors
            x = up.upload(host_clips)              # device view, valid until the next-but-one upload()
            ... launch work that reads x on the current stream ...
            up.release()                           # call after enqueuing the readers of x
    """

    def __init__(self, batch_shape, device, dtype=torch.uint8, crop=None):
        """crop = (y0, x0, h, w): upload only that window of every frame (the columns/rows the network's center crop
        keeps) with a strided 2D DMA — 44 % fewer PCIe bytes for 224x398 frames cropped to 224x224.  The returned
        device tensor then has shape (B, T, 3, h, w) and must be run with crop offsets (0, 0)."""
        self.stream = torch.cuda.Stream(device=device)
        self.crop = crop
        if crop is not None:
            batch_shape = tuple(batch_shape[:-2]) + (crop[2], crop[3])
        self.bufs = [torch.empty(batch_shape, dtype=dtype, device=device) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self._used = [False, False]

    def upload(self, host_clips):
        k = self.i & 1
        n = host_clips.shape[0]
        with torch.cuda.stream(self.stream):
            if self._used[k]:
                self.stream.wait_event(self.free[k])         # readers of this buffer (two uploads ago) are done
            if self.crop is None:
                self.bufs[k][:n].copy_(host_clips, non_blocking=True)
            else:
                assert host_clips.is_contiguous() and host_clips.is_pinned() and host_clips.dtype == torch.uint8
                y0, x0, h, w = self.crop
                in_h, in_w = host_clips.shape[-2:]
                planes = host_clips.numel() // (in_h * in_w)
                dst = self.bufs[k][:n]
                if h == in_h:           # rows of all planes are equally spaced: one 2D copy for the whole batch
                    _memcpy2d_async(dst.data_ptr(), w, host_clips.data_ptr() + x0, in_w, w, planes * in_h, self.stream.cuda_stream)
                else:                   # vertical crop too: one 2D copy per plane would be too many calls -> 3D not exposed; fall back
                    dst.copy_(host_clips[..., y0:y0 + h, x0:x0 + w], non_blocking=True)
            self.ready[k].record(self.stream)
        torch.cuda.current_stream().wait_event(self.ready[k])
        self._k = k
        return self.bufs[k][:n]

    def release(self):
        self.free[self._k].record(torch.cuda.current_stream())
        self._used[self._k] = True
        self.i += 1


def nms_events(ev, k, window, threshold, soft):
    """(soft-)NMS of the high-recall events of one video; returns numpy (frame i32, label i32, score f64)."""
    of, ol, os_, oc = ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], k, window, threshold, soft)
    n = int(oc.item())            # the only device->host sync of the post-processing
    return of[:n].cpu().numpy(), ol[:n].cpu().numpy(), os_[:n].cpu().numpy()


class PendingEvents:
    """(soft-)NMS result whose device->host copy is still in flight: lets the caller keep launching the next
    video instead of synchronising per video like util/eval.py does.  `.get()` blocks and returns numpy arrays."""

    def __init__(self, ev, k, window, threshold, soft):
        of, ol, os_, oc = ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], k, window, threshold, soft)
        self._host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (of, ol, os_, oc)]
        for h, d in zip(self._host, (of, ol, os_, oc)):
            h.copy_(d, non_blocking=True)
        self._dev = (of, ol, os_, oc)          # keep alive until the copies have run
        self._done = torch.cuda.Event()
        self._done.record()
        self.nbytes = sum(h.numel() * h.element_size() for h in self._host)

    def get(self):
        self._done.synchronize()
        self._dev = None
        n = int(self._host[3][0])
        return self._host[0][:n].numpy(), self._host[1][:n].numpy(), self._host[2][:n].numpy()


def events_to_dicts(video, fps, frames, labels, scores, classes_inv):
    """The reference's wire format: {'video', 'events': [{'label','frame','score'}], 'fps'}."""
    return {'video': video, 'fps': fps,
            'events': [{'label': classes_inv[int(l)], 'frame': int(f), 'score': float(s)}
                       for f, l, s in zip(frames, labels, scores)]}


# ---------------------------------------------------------------------------------------------------------------------
# video-level inference: every unique frame passes the clip-independent layers once
# ---------------------------------------------------------------------------------------------------------------------
class VideoInference:
    """Streams the frames of a list of videos through the engine and returns per-video device-resident scores.

    The reference slices a video into clips that overlap by 75 % (dataset/frame.py:409-423) and runs each clip through the
    whole network (util/eval.py:289-349), so stem + s1 + s2 — which see one frame at a time; the first cross-frame
    operator is the GatedShift of s3.b1 (model/shift.py:47-59) — process every frame four times and every frame crosses
    PCIe four times.  Here the frames of all videos form ONE stream: unique frames are uploaded once into one of two
    `frames_per_chunk`-frame device buffers (side stream, overlapping the kernels of the previous chunk), run through
    InferenceEngine.lower() once (twice with the flipped TTA view) and filed into a ring of per-frame features in HBM; clip
    batches are gathered from the ring (tdeed_gather_rows), finished by upper(), and the temporal stack + heads run over the
    pooled features of a few batches at once.  Frames before 0 /
    past the end of a video read the features of the all-zero frame, exactly what the reference's zero padding
    (dataset/frame.py:622-625) produces.  Results are bit-identical to per-clip execution (tests/test_gpu_video.py): no
    kernel of the lower part mixes frames, and clip order == accumulation order.

        vi = VideoInference(engine, in_hw=(224, 398), clips_per_batch=57, frames_per_chunk=1425, flips=(False,))
        scores = vi.run(videos, pieces)   # videos: [(name, video_len, [clip starts])]; pieces: iterator of pinned uint8
                                          # (n, 3, H, W) host tensors (any n) = the videos' frames back to back
    """

    def __init__(self, engine, in_hw, clips_per_batch=57, frames_per_chunk=None, flips=(False,), clip_len=None,
                 upload_crop=True, temporal_batches=None):
        self.eng = engine
        self.dev = engine.device
        self.T = clip_len or engine.cfg.clip_len
        self.B = clips_per_batch
        # The temporal stack + heads run once per `temporal_batches` backbone batches (their pooled features wait in a stash):
        # ~40 small launches whose time barely depends on the number of clips (3 x 57 clips: 3.5 -> ~2 ms per video).
        self.temporal_batches = temporal_batches or max(1, min(8, 171 // self.B))
        self.N = frames_per_chunk or max(self.T, self.B * max(1, self.T // 4))
        self.flips = tuple(flips)
        self.K = engine.cfg.num_classes + 1
        in_h, in_w = in_hw
        cy, cx, ch, cw = engine.crop_window(in_h, in_w)
        # only the window the network's center crop keeps crosses PCIe (strided 2D DMA); rows must stay whole for that
        self.upload_crop = (cy, cx, ch, cw) if (upload_crop and ch == in_h and cw != in_w) else None
        self.dev_crop = (0, 0, ch, cw) if self.upload_crop else (cy, cx, ch, cw)
        self.in_hw = (in_h, in_w)
        self.dev_hw = (ch, cw) if self.upload_crop else (in_h, in_w)
        self.W_slots = self.B * self.T + 2 * self.N
        self.ring = None            # per flip: (W_slots, h, w, c)
        self.pad = None             # per flip: (h, w, c) features of the all-zero frame
        self.stream = torch.cuda.Stream(device=self.dev)
        self.bufs = [torch.zeros((self.N, 3) + self.dev_hw, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self._used = [False, False]
        self._k, self._fill = 0, 0
        self._inflight = []         # (event, host pieces): keeps pinned pieces alive until their (raw 2D) DMA has run
        self._pieces = []
        self.h2d_bytes = 0
        self.frames_in = 0          # frames whose features are in the ring (global stream index of the next one)
        self.clips_out = 0
        self.use_graphs = True      # False: launch lower / upper eagerly (per-kernel profiling with engine.prof)
        self.launches = 0           # kernels launched outside the engine (gather / accumulate)

    def _prepare(self):
        eng = self.eng
        if self.ring is not None and self._version == eng.version:
            return
        black = torch.zeros((1, 3) + self.dev_hw, dtype=torch.uint8, device=self.dev)
        pad = [eng.lower(black, flip=flip, crop=self.dev_crop)[0].clone() for flip in self.flips]
        if self.ring is None:
            self.ring = [torch.empty((self.W_slots,) + tuple(f.shape), dtype=f.dtype, device=self.dev) for f in pad]
            self.xg = [torch.empty((self.B * self.T,) + tuple(f.shape), dtype=f.dtype, device=self.dev) for f in pad]
            self.stash = [torch.empty((self.temporal_batches * self.B * self.T, eng.cfg.feat_dim), dtype=torch.float32, device=self.dev)
                          for _ in pad]
            self.pad = pad
        else:                       # new weights (engine.load_state): same buffers (graphs may be keyed on them), new contents
            for dst, f in zip(self.pad, pad):
                dst.copy_(f)
        self._version = eng.version

    def _feed(self, piece):
        """Queue the H2D copy of a pinned host piece (n,3,H,W) into the filling device buffer; flush full buffers."""
        assert piece.dtype == torch.uint8 and piece.dim() == 4 and tuple(piece.shape[-2:]) == self.in_hw, \
            'expected uint8 (n,3,%d,%d) frames, got %s %s' % (self.in_hw + (piece.dtype, tuple(piece.shape)))
        if not piece.is_cuda and not piece.is_pinned():
            piece = piece.contiguous().pin_memory()
        off, n = 0, piece.shape[0]
        while off < n:
            take = min(self.N - self._fill, n - off)
            k = self._k
            if piece.is_cuda and self._fill == 0 and take == self.N and piece.is_contiguous():
                # a whole chunk already resident in HBM: lower() reads it in place (its own crop window), no staging copy
                self._flush(direct=piece[off:off + take])
                off += take
                continue
            with torch.cuda.stream(self.stream):
                if self._fill == 0 and self._used[k]:
                    self.stream.wait_event(self.free[k])          # the kernels that read this buffer two chunks ago are done
                dst = self.bufs[k][self._fill:self._fill + take]
                src = piece[off:off + take]
                if piece.is_cuda:                                  # frames already resident in HBM (bench `value` arm)
                    self.stream.wait_stream(torch.cuda.current_stream())
                    if self.upload_crop is None:
                        dst.copy_(src)
                    else:
                        dst.copy_(src[..., self.upload_crop[1]:self.upload_crop[1] + self.upload_crop[3]])
                elif self.upload_crop is None:
                    dst.copy_(src, non_blocking=True)
                else:
                    _, x0, h, w = self.upload_crop
                    in_h, in_w = self.in_hw
                    _memcpy2d_async(dst.data_ptr(), w, src.data_ptr() + x0, in_w, w, take * 3 * in_h, self.stream.cuda_stream)
            if not piece.is_cuda:
                self.h2d_bytes += take * 3 * self.dev_hw[0] * self.dev_hw[1]
            self._fill += take
            off += take
            self._pieces.append(piece)
            if self._fill == self.N:
                self._flush()

    def _flush(self, direct=None):
        """Run lower() on the filling buffer (the graph always processes N frames; stale tail rows are ignored) and file
        the features of its `_fill` valid frames into the ring(s).  direct: a resident (N,3,H,W) device chunk to read in
        place instead of the staging buffer."""
        n, k = (self.N, None) if direct is not None else (self._fill, self._k)
        if n == 0:
            return
        if direct is None:
            self.ready[k].record(self.stream)
            self._inflight = [(e, p) for e, p in self._inflight if not e.query()]
            self._inflight.append((self.ready[k], self._pieces))
            self._pieces = []
            torch.cuda.current_stream().wait_event(self.ready[k])
            src, crop = self.bufs[k], self.dev_crop
        else:
            src, crop = direct, self.eng.crop_window(*self.in_hw)
        for fi, flip in enumerate(self.flips):
            lower = self.eng.lower_graphed if self.use_graphs else self.eng.lower
            feat = lower(src, flip=flip, crop=crop)
            ops.scatter_rows_ring(feat, n, self.ring[fi], self.frames_in % self.W_slots)
            self.launches += 1
        if direct is None:
            self.free[k].record(torch.cuda.current_stream())
            self._used[k] = True
            self._k, self._fill = k ^ 1, 0
        self.frames_in += n

    def run(self, videos, chunks, on_video=None):
        """videos: [(name, video_len, starts)] in stream order; chunks: iterator of pinned uint8 (n,3,H,W) host pieces.
        Returns {name: VideoScores}.  on_video(name, VideoScores) is called as soon as a video's last clip has been
        accumulated (lets the caller start event extraction / D2H while later videos still compute)."""
        self._prepare()
        eng, T, B, K = self.eng, self.T, self.B, self.K
        tta = len(self.flips) > 1
        base, g = {}, 0
        for name, vlen, _ in videos:
            base[name] = g
            g += vlen
        total_frames = g
        clips = [(name, vlen, s) for name, vlen, starts in videos for s in starts]
        last_clip = {}
        for ci, (name, _, _) in enumerate(clips):
            last_clip[name] = ci
        scores = {name: VideoScores(vlen, K, self.dev) for name, vlen, _ in videos}
        chunks = iter(chunks)
        self.frames_in, self._fill = 0, 0
        pending = []                                     # batches whose features wait in the stash for the temporal stack
        for lo in range(0, len(clips), B):
            batch = clips[lo:lo + B]
            need = 0
            first, los, his = [0] * B, [0] * B, [0] * B        # clips beyond len(batch) (ragged last batch): all padding
            for bi, (name, vlen, s) in enumerate(batch):
                lo_t, hi_t = min(T, max(0, -s)), max(0, min(T, vlen - s))
                if hi_t > lo_t:
                    first[bi], los[bi], his[bi] = (base[name] + s) % self.W_slots, lo_t, hi_t
                    need = max(need, base[name] + s + hi_t)
            while self.frames_in < min(need, total_frames):
                piece = next(chunks, None)
                if piece is not None:
                    self._feed(piece)
                elif self._fill:
                    self._flush()                # end of the stream: ragged last chunk
                else:
                    raise RuntimeError('frame stream ended after %d frames; the clip list needs %d' % (self.frames_in, need))
            # backbone (stages 3-4) of this batch -> pooled features, parked in slot len(pending) of the feature stash
            j = len(pending)
            for fi in range(len(self.flips)):
                ops.gather_clip_rows(self.ring[fi], self.pad[fi], self.xg[fi], T, first, los, his)
                feat = eng.upper_feat_graphed(self.xg[fi], B, T) if self.use_graphs else eng.upper(self.xg[fi], B, T)
                self.stash[fi][j * B * T:(j + 1) * B * T].copy_(feat)
                self.launches += 2
            pending.append((lo, batch))
            if len(pending) < self.temporal_batches and lo + B < len(clips):
                continue
            # temporal stack + heads once for all parked batches (clip results do not depend on the batch they ride in)
            k = len(pending)
            outs = []
            for fi in range(len(self.flips)):
                fview = self.stash[fi][:k * B * T]
                if self.use_graphs:
                    _, _, probs = eng.temporal_heads_graphed(fview, k * B, T)
                else:
                    _, _, probs = eng.heads(eng.temporal(fview.view(k * B, T, eng.cfg.feat_dim)))
                outs.append(probs)
            rep = len(outs) if tta else 1
            for j, (blo, bbatch) in enumerate(pending):
                nb = len(bbatch)
                if tta:  # the reference adds plain then flipped view clip by clip (util/eval.py:321-349): keep that order
                    probs = torch.stack([o[j * B:j * B + nb] for o in outs], dim=1).reshape(nb * len(outs), T, K)
                else:
                    probs = outs[0][j * B:j * B + nb]
                bi = 0
                while bi < nb:                      # clips of one batch may belong to several videos
                    name = bbatch[bi][0]
                    bj = bi
                    while bj < nb and bbatch[bj][0] == name:
                        bj += 1
                    st = [bbatch[i][2] for i in range(bi, bj) for _ in range(rep)]
                    scores[name].add(probs[bi * rep:bj * rep], st, tta=tta)
                    self.launches += 1
                    if on_video is not None and last_clip[name] == blo + bj - 1:
                        on_video(name, scores[name])
                    bi = bj
                self.clips_out += nb
            pending = []
        return scores


class ThreadedFrameSource:
    """Ordered iterator over `pieces[k]` (uint8 (n,3,H,W) host tensors) produced by worker THREADS straight into a ring of
    pinned buffers: no process boundary, no pickling / shared-memory transport, no separate pin-memory thread — the decode
    (torchvision.io releases the GIL) and the copy into pinned memory run in parallel, the consumer hands the pinned slot to
    the DMA engine.  A torch DataLoader moved the same bytes at ~4 GB/s (worker process -> shm -> pin thread), below what one
    B200 consumes.  A slot is recycled only after the upload that read it has completed (event on `stream`)."""

    def __init__(self, pieces, stream, workers=8, depth=24, slots=None):
        """slots: a list (kept by the caller across runs) that holds the pinned buffers, so that they are allocated once
        (cudaHostAlloc of a 400 MB ring costs more than a whole video's inference)."""
        import threading
        self.pieces, self.stream, self.n = pieces, stream, len(pieces)
        self.depth = max(depth, workers + 2)
        if slots is not None and len(slots) < self.depth:
            slots.extend([None] * (self.depth - len(slots)))
        self._slots = slots if slots is not None else [None] * self.depth
        self._events = [None] * self.depth
        self._done = {}
        self._cv = threading.Condition()
        self._next_task = 0
        self._consumed = 0            # pieces handed to the consumer so far
        self._err = None
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(min(workers, max(1, self.n)))]
        for t in self._threads:
            t.start()

    def _work(self):
        try:
            while True:
                with self._cv:
                    while self._next_task < self.n and self._next_task >= self._consumed + self.depth and self._err is None:
                        self._cv.wait()           # the ring is full: wait for the consumer
                    if self._next_task >= self.n or self._err is not None:
                        return
                    k = self._next_task
                    self._next_task += 1
                    ev = self._events[k % self.depth]
                if ev is not None:
                    ev.synchronize()              # the upload that read this slot (depth pieces ago) has finished
                piece = self.pieces[k]
                slot = self._slots[k % self.depth]
                if slot is None or slot.shape[1:] != piece.shape[1:] or slot.shape[0] < piece.shape[0]:
                    slot = torch.empty(piece.shape, dtype=piece.dtype, pin_memory=True)
                    self._slots[k % self.depth] = slot
                view = slot[:piece.shape[0]]
                view.copy_(piece)
                with self._cv:
                    self._done[k] = view
                    self._cv.notify_all()
        except BaseException as exc:                # surface worker errors in the consumer
            with self._cv:
                self._err = exc
                self._cv.notify_all()

    def __len__(self):
        return self.n

    def __iter__(self):
        for k in range(self.n):
            with self._cv:
                while k not in self._done and self._err is None:
                    self._cv.wait()
                if self._err is not None:
                    raise self._err
                view = self._done.pop(k)
            yield view
            # the consumer has enqueued its copy of `view` on self.stream by the time it asks for the next piece
            ev = torch.cuda.Event()
            ev.record(self.stream)
            with self._cv:
                self._events[k % self.depth] = ev
                self._consumed = k + 1
                self._cv.notify_all()
