"""TEST INFRASTRUCTURE — CPU oracle (numpy + plain Python) for T-DEED's per-video post-processing.

Restates, on flat arrays instead of lists of dicts, what the reference's util/eval.py does
between `model.predict` and the stored event list:
  clip_starts               dataset/frame.py:409-423,451
  accumulate_batched / _tta util/eval.py:299-349
  frame_predictions         util/eval.py:87-140 (and _challenge :142-193)
  nms                       util/eval.py:195-227
  soft_nms                  util/eval.py:229-261
Pinned against the unmodified reference functions in tests/test_oracle_vs_reference.py (run in the
build container) and by tests/golden/postproc_*.npz generated from the reference.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np


def clip_starts(num_frames, clip_len=100, overlap_len=75, stride=1, pad_len=5):
    """Start frames (already divided by stride) of the sliding inference clips of one video."""
    return [s // stride for s in range(-pad_len * stride,
                                       max(0, num_frames - overlap_len * stride),
                                       (clip_len - overlap_len) * stride)]


def video_len(num_frames, stride=1):
    """dataset/frame.py:489-497: ceil(num_frames / stride)."""
    return int(np.ceil(num_frames / stride))


def accumulate_batched(scores, support, pred, start):
    """One clip of the batched path (util/eval.py:303-317).  pred: (T,K) float32."""
    if start < 0:
        pred = pred[-start:, :]
        start = 0
    end = start + pred.shape[0]
    if end >= scores.shape[0]:
        end = scores.shape[0]
        pred = pred[:end - start, :]
    scores[start:end, :] += pred
    support[start:end] += (pred.sum(axis=1) != 0) * 1


def accumulate_tta(scores, support, pred, start):
    """One view of the TTA path (util/eval.py:321-349).  pred: (1,T,K)."""
    if start < 0:
        pred = pred[:, -start:, :]
        start = 0
    end = start + pred.shape[1]
    if end >= scores.shape[0]:
        end = scores.shape[0]
        pred = pred[:, :end - start, :]
    scores[start:end, :] += np.sum(pred, axis=0)
    support[start:end] += pred.shape[0]


def frame_predictions(scores, support, threshold=0.01):
    """Normalise in place and extract events of ONE video.

    Returns (pred int64 (L,), events, events_high_recall) with each event list a tuple of arrays
    (frame int32, label int32 [class index 1..K-1], score float32), frame-major then class order.
    """
    if np.min(support) == 0:
        support[support == 0] = 1
    scores /= support[:, None]
    pred = np.argmax(scores, axis=1)
    fr = np.nonzero(pred != 0)[0]
    events = (fr.astype(np.int32), pred[fr].astype(np.int32), scores[fr, pred[fr]].astype(np.float32))
    thr = np.float32(threshold)              # numpy-2 weak scalar: compared in float32
    fi, ci = np.nonzero(scores[:, 1:] >= thr)
    hr = (fi.astype(np.int32), (ci + 1).astype(np.int32), scores[fi, ci + 1].astype(np.float32))
    return pred, events, hr


def _by_label(frames, labels):
    """Label buckets in first-appearance order (dict insertion order in the reference)."""
    order, seen = [], set()
    for l in labels.tolist():
        if l not in seen:
            seen.add(l)
            order.append(l)
    return [(l, np.nonzero(labels == l)[0]) for l in order]


def nms(frames, labels, scores, window, threshold=0.0):
    """Greedy hard NMS of one video's events -> (frames, labels, scores) sorted by frame (stable)."""
    out = []
    for bi, (l, idx) in enumerate(_by_label(frames, labels)):
        w = window[bi] if isinstance(window, list) else window
        f = frames[idx].astype(np.int64)
        s = scores[idx].astype(np.float64)           # .item() -> Python float
        alive = np.ones(len(idx), bool)
        while alive.any():
            cand = np.nonzero(alive)[0]
            j = cand[np.argmax(s[cand])]             # first maximum = lowest frame on ties
            if s[j] < threshold:
                break
            out.append((int(f[j]), int(l), float(s[j])))
            alive &= ~(np.abs(f - f[j]) <= w)
    return _sorted(out, scores.dtype)


def soft_nms(frames, labels, scores, window, threshold=0.01):
    """Soft NMS (quadratic decay inside +-window) -> (frames, labels, float64 scores) sorted by frame."""
    out = []
    for bi, (l, idx) in enumerate(_by_label(frames, labels)):
        w = window[bi] if isinstance(window, list) else window
        f = frames[idx].astype(np.int64)
        s = scores[idx].astype(np.float64)
        alive = np.ones(len(idx), bool)
        while alive.any():
            cand = np.nonzero(alive)[0]
            j = cand[np.argmax(s[cand])]
            if s[j] < threshold:
                break
            out.append((int(f[j]), int(l), float(s[j])))
            near = alive & (np.abs(f - f[j]) <= w)
            d = np.abs(f[j] - f[near])
            s[near] = s[near] * (d ** 2) / (w ** 2)
            alive[j] = False
    return _sorted(out, np.float64)


def _sorted(out, score_dtype):
    out.sort(key=lambda e: e[0])                       # stable, like list.sort(key=frame)
    if not out:
        return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, score_dtype)
    fr, lb, sc = zip(*out)
    return np.asarray(fr, np.int32), np.asarray(lb, np.int32), np.asarray(sc, score_dtype)


# ---- adapters between the array form and the reference's list-of-dicts form (used by tests) ----

def to_dicts(video, frames, labels, scores, classes_inv, fps=25.0):
    return {'video': video, 'fps': fps,
            'events': [{'label': classes_inv[int(l)], 'frame': int(f), 'score': float(s)}
                       for f, l, s in zip(frames, labels, scores)]}


def from_dicts(video_pred, classes):
    ev = video_pred['events']
    return (np.asarray([e['frame'] for e in ev], np.int32),
            np.asarray([classes[e['label']] for e in ev], np.int32),
            np.asarray([e['score'] for e in ev], np.float64))
