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
        st = torch.as_tensor(np.asarray(starts, np.int32)).to(self.scores.device, non_blocking=True)
        ops.clip_accumulate(self.scores, self.support, probs.contiguous(), st, 1 if tta else 0)

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
