"""TEST INFRASTRUCTURE — synthetic stand-ins for the reference's inference dataset and for a model, used to drive
`util.eval.evaluate` (the reference's and this repo's) on identical inputs.

SyntheticVideoDataset mirrors the attribute surface of dataset/frame.py:385-517 (ActionSpotVideoDataset) that
util/eval.py touches — `_dataset`, `_stride`, `_clip_len`, `_clips`, `_labels`, `_frame_reader.load_frames`, `videos`,
`labels`, `get_labels`, `__getitem__` -> {'video', 'start', 'frame'} — with frames generated on the fly (or read from
a directory of JPEGs written by `write_jpegs`) instead of an extracted-frames tree.

CannedModel.predict returns per-clip probabilities looked up by the (video id, frame index) tag that the synthetic
frames carry in their first pixels, so the reference's evaluate() loop can be replayed with known inputs.
"""
import copy
import math
import os

import numpy as np
import torch

DEFAULT_PAD_LEN = 5


def frame_pixels(video_id, index, hw, seed=0):
    """Deterministic uint8 (3,H,W) frame; pixels [0, 0, 0:3] carry (video_id, index // 256, index % 256)."""
    h, w = hw
    rng = np.random.default_rng([seed, video_id, index])
    img = rng.integers(0, 256, size=(3, h, w), dtype=np.uint8)
    img[0, 0, 0], img[0, 0, 1], img[0, 0, 2] = video_id, index // 256, index % 256
    img[1, 0, 0] = 255                                   # marks "not a padding frame"
    return torch.from_numpy(img)


class SyntheticFrameReader:
    """load_frames with the semantics of dataset/frame.py:546-626 (FrameReaderVideo): frames before 0 are counted as
    start padding, frames past the end are 'missing files' (padded only if pad=True), -1 when nothing could be read."""

    def __init__(self, lengths, hw, video_ids, seed=0, jpeg_dir=None):
        self.lengths, self.hw, self.video_ids, self.seed, self.jpeg_dir = lengths, hw, video_ids, seed, jpeg_dir

    def read(self, video, index):
        if self.jpeg_dir is not None:
            import torchvision
            return torchvision.io.read_image(os.path.join(self.jpeg_dir, video, 'frame%d.jpg' % index))
        return frame_pixels(self.video_ids[video], index, self.hw, self.seed)

    def load_frames(self, video_name, start, end, pad=False, stride=1, source_info=None):
        ret, n_pad_start, n_pad_end = [], 0, 0
        for frame_num in range(start, end, stride):
            if frame_num < 0:
                n_pad_start += 1
                continue
            if frame_num >= self.lengths[video_name]:
                n_pad_end += 1
                continue
            ret.append(self.read(video_name, frame_num))
        if len(ret) == 0:
            return -1
        ret = torch.stack(ret, dim=0)
        if n_pad_start > 0 or (pad and n_pad_end > 0):
            ret = torch.nn.functional.pad(ret, (0, 0, 0, 0, 0, 0, n_pad_start, n_pad_end if pad else 0))
        return ret


class SyntheticVideoDataset(torch.utils.data.Dataset):

    def __init__(self, classes, lengths, hw, clip_len, overlap_len, stride=1, pad_len=DEFAULT_PAD_LEN, dataset='fs_comp',
                 fps=25.0, seed=0, events_per_100=3.0, jpeg_dir=None):
        names = sorted(lengths)
        rng = np.random.default_rng(seed + 1)
        inv = sorted(classes.values())
        self._labels = []
        for name in names[::-1]:                         # label-file order != sorted order, like real label files
            n = lengths[name]
            count = max(1, int(n * events_per_100 / 100))
            frames = sorted(set(int(f) for f in rng.integers(0, n, size=count)))
            cls_names = {v: k for k, v in classes.items()}
            self._labels.append({'video': name, 'num_frames': n, 'fps': fps,
                                 'events': [{'frame': f, 'label': cls_names[int(rng.choice(inv))]} for f in frames]})
        self._class_dict = classes
        self._video_idxs = {x['video']: i for i, x in enumerate(self._labels)}
        self._clip_len, self._stride, self._dataset = clip_len, stride, dataset
        self._frame_reader = SyntheticFrameReader(lengths, hw, {n: i for i, n in enumerate(names)}, seed, jpeg_dir)
        self._clips = []
        for l in self._labels:
            for i in range(-pad_len * stride, max(0, l['num_frames'] - overlap_len * stride), (clip_len - overlap_len) * stride):
                self._clips.append((l['video'], i))

    def __len__(self):
        return len(self._clips)

    def __getitem__(self, idx):
        video_name, start = self._clips[idx]
        frames = self._frame_reader.load_frames(video_name, start, start + self._clip_len * self._stride, pad=True, stride=self._stride)
        return {'video': video_name, 'start': start // self._stride, 'frame': frames}

    def get_labels(self, video):
        meta = self._labels[self._video_idxs[video]]
        labels = np.zeros(math.ceil(meta['num_frames'] / self._stride), np.int64)
        for event in meta['events']:
            if event['frame'] < meta['num_frames']:
                labels[event['frame'] // self._stride] = self._class_dict[event['label']]
        return labels

    @property
    def videos(self):
        return sorted([(v['video'], math.ceil(v['num_frames'] / self._stride), v['fps'] / self._stride) for v in self._labels])

    @property
    def labels(self):
        if self._stride == 1:
            return self._labels
        out = []
        for x in self._labels:
            y = copy.deepcopy(x)
            y['fps'] /= self._stride
            y['num_frames'] //= self._stride
            for e in y['events']:
                e['frame'] //= self._stride
            out.append(y)
        return out


def write_jpegs(dataset, root):
    """Materialise the synthetic frames as <root>/<video>/frame<i>.jpg (the reference's fs_comp layout) and switch the
    reader to them, so that the JPEG decode path is exercised too."""
    import torchvision
    r = dataset._frame_reader
    for video, n in r.lengths.items():
        os.makedirs(os.path.join(root, video), exist_ok=True)
        for i in range(n):
            torchvision.io.write_jpeg(frame_pixels(r.video_ids[video], i, r.hw, r.seed), os.path.join(root, video, 'frame%d.jpg' % i),
                                      quality=90)
    r.jpeg_dir = root


class CannedModel:
    """predict() returns canned probabilities: probs[video_id][flip][start_of_clip] (T,K), found through the frame tags."""

    def __init__(self, dataset, num_classes_p1, seed=0, zero_rows=0.1, smooth=5, temp=2.0):
        self.clip_len, self.stride, self.k = dataset._clip_len, dataset._stride, num_classes_p1
        ids = dataset._frame_reader.video_ids
        rng = np.random.default_rng(seed + 2)
        self.table = {}
        for video, start in dataset._clips:
            for flip in (False, True):
                p = synth_scores(rng, self.clip_len, num_classes_p1, smooth, temp)
                p[rng.random(self.clip_len) < zero_rows] = 0          # frames never hit by the displacement scatter
                self.table[(ids[video], start // self.stride, flip)] = p

    def _key(self, clip, flip):
        first = int(torch.nonzero(clip[:, 1, 0, 0] == 255)[0])          # first non-padding frame
        vid = int(clip[first, 0, 0, 0])
        index = int(clip[first, 0, 0, 1]) * 256 + int(clip[first, 0, 0, 2])
        return (vid, index // self.stride - first, flip)

    def predict(self, seq, use_amp=True, augment_inference=False):
        seq = torch.as_tensor(seq)
        if seq.dim() == 4:
            seq = seq.unsqueeze(0)
        probs = np.stack([self.table[self._key(c, bool(augment_inference))] for c in seq])
        return np.argmax(probs, axis=2), probs


def synth_scores(rng, length, k, smooth=5, temp=2.0):
    """Smoothed softmax(N(0, temp^2)) so that a few % of (frame, class) cells exceed 0.01."""
    z = rng.normal(0, temp, size=(length + smooth - 1, k)).astype(np.float32)
    z[:, 0] += 3.0
    kern = np.ones(smooth, np.float32) / smooth
    z = np.stack([np.convolve(z[:, j], kern, mode='valid') for j in range(k)], axis=1) * np.float32(2.5)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
