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


def nms_events(ev, k, window, threshold, soft):
    """(soft-)NMS of the high-recall events of one video; returns numpy (frame i32, label i32, score f64)."""
    of, ol, os_, oc = ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], k, window, threshold, soft)
    n = int(oc.item())            # the only device->host sync of the post-processing
    return of[:n].cpu().numpy(), ol[:n].cpu().numpy(), os_[:n].cpu().numpy()


def events_to_dicts(video, fps, frames, labels, scores, classes_inv):
    """The reference's wire format: {'video', 'events': [{'label','frame','score'}], 'fps'}."""
    return {'video': video, 'fps': fps,
            'events': [{'label': classes_inv[int(l)], 'frame': int(f), 'score': float(s)}
                       for f, l, s in zip(frames, labels, scores)]}
