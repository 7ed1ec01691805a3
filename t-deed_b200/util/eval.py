"""Drop-in for the reference's util/eval.py: same names, signatures and list-of-dict wire format, so that
`from util.eval import evaluate, valMAP_SN, evaluate_SNB` (train_tdeed.py:24, evaluate_tdeed_challenge.py:22) resolves here
when `t-deed_b200/` is first on sys.path.

Hot-path half — built here on the sm_100a kernels of libtdeed_sm100 (no CPU fallback):
    evaluate                               util/eval.py:264-419   video-level engine + device-resident post-processing
    process_frame_predictions[_challenge]  util/eval.py:87-193    tdeed_extract_events
    non_maximum_supression                 util/eval.py:195-227   tdeed_nms (hard)
    soft_non_maximum_supression            util/eval.py:229-261   tdeed_nms (soft, fp64)
    ErrorStat / ForegroundF1               util/eval.py:34-85     vectorised confusion-matrix counters (same numbers)
Not on the hot path — `valMAP_SN`, `evaluate_SNB`, `aux_evaluate`, `label2vector`, `predictions2vector`
(util/eval.py:422-674, SoccerNet evaluator glue): NOT retyped.  They are served by the module-level __getattr__ below,
which loads the reference checkout's own util/eval.py (the next one on sys.path) under a private name and re-exports
those five names.

`t-deed_b200/util/` deliberately has NO __init__.py: like the reference's `util/` it is a namespace-package portion; this
portion comes first on sys.path, so `util.eval`, `util.score`, `util.io`, `util.dataset` resolve here and anything else in
the reference's `util/` stays reachable.
"""
import importlib.util
import math
import os
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from tqdm import tqdm

from tdeed_b200 import ops
from tdeed_b200.parallel import gather_video_results, shard_videos, world
from tdeed_b200.pipeline import PendingEvents, ThreadedFrameSource, VideoInference, VideoScores

# Constants (util/eval.py:24-32 of the reference)
TOLERANCES = [1, 2, 4]
WINDOWS = [1, 3]
TOLERANCES_SN = [3, 6]
WINDOWS_SN = [3, 6]
TOLERANCES_SNB = [6, 12]
WINDOWS_SNB = [6, 12]
WINDOWS_T = [1, 3]
WINDOWS_FG = [1, 3]
INFERENCE_BATCH_SIZE = 4

# video-level engine tuning: clips per upper batch at 224 x 224 (scaled down with the frame area), host piece size
STREAM_CLIPS_PER_BATCH = 57
STREAM_PIECE_FRAMES = 64
STREAM_WORKERS = 4 * 2

# ---------------------------------------------------------------------------------------------------------------------
# the non-hot-path half lives in the reference checkout: re-export, do not retype
# ---------------------------------------------------------------------------------------------------------------------
_REFERENCE_NAMES = ('valMAP_SN', 'evaluate_SNB', 'aux_evaluate', 'label2vector', 'predictions2vector')
_reference_eval = None


def _load_reference_eval():
    global _reference_eval
    if _reference_eval is None:
        here = os.path.dirname(os.path.abspath(__file__))
        for entry in sys.path:
            cand = os.path.join(entry or '.', 'util', 'eval.py')
            if os.path.isfile(cand) and os.path.dirname(os.path.abspath(cand)) != here:
                spec = importlib.util.spec_from_file_location('_tdeed_reference_util_eval', cand)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _reference_eval = mod
                break
        else:
            raise ImportError('util.eval: %s are served by the T-DEED checkout\'s own util/eval.py (SoccerNet evaluator glue, '
                              'off the hot path) — put the checkout on sys.path after t-deed_b200/' % (_REFERENCE_NAMES,))
    return _reference_eval


def __getattr__(name):
    if name in _REFERENCE_NAMES:
        return getattr(_load_reference_eval(), name)
    raise AttributeError('module %r has no attribute %r' % (__name__, name))


# ---------------------------------------------------------------------------------------------------------------------
# frame-level statistics (util/eval.py:34-85): one confusion matrix instead of per-frame Python dict updates
# ---------------------------------------------------------------------------------------------------------------------
class ErrorStat:
    """Frame-level error rate."""

    def __init__(self):
        self._total = 0
        self._err = 0

    def update(self, true, pred):
        true, pred = np.asarray(true), np.asarray(pred)
        self._err += int(np.count_nonzero(true != pred))
        self._total += int(true.shape[0])

    def get(self):
        return self._err / self._total

    def get_acc(self):
        return 1. - self.get()


class ForegroundF1:
    """Exact-frame F1 per class and for 'any foreground' (key None).  Counts come from a confusion matrix
    C[true, pred]:  any: tp = C[1:,1:].sum, fp = C[0,1:].sum, fn = C[1:,0].sum;  class k: tp = C[k,k],
    fp = C[:,k].sum - C[k,k], fn = C[k,:].sum - C[k,k] — the same numbers the reference's per-frame updates produce."""

    def __init__(self):
        self._c = np.zeros((1, 1), np.int64)

    def _grow(self, n):
        if n > self._c.shape[0]:
            c = np.zeros((n, n), np.int64)
            c[:self._c.shape[0], :self._c.shape[1]] = self._c
            self._c = c

    def update(self, true, pred):
        """Scalars (reference call style) or equal-length integer arrays."""
        true = np.atleast_1d(np.asarray(true, np.int64))
        pred = np.atleast_1d(np.asarray(pred, np.int64))
        self._grow(int(max(true.max(initial=0), pred.max(initial=0))) + 1)
        np.add.at(self._c, (true, pred), 1)

    def merge_counts(self, counts):
        counts = np.asarray(counts, np.int64)
        self._grow(counts.shape[0])
        self._c[:counts.shape[0], :counts.shape[1]] += counts

    def tp_fp_fn(self, k):
        c = self._c
        if k is None:
            return int(c[1:, 1:].sum()), int(c[0, 1:].sum()), int(c[1:, 0].sum())
        if k >= c.shape[0]:
            return 0, 0, 0
        return int(c[k, k]), int(c[:, k].sum() - c[k, k]), int(c[k, :].sum() - c[k, k])

    def get(self, k):
        tp, fp, fn = self.tp_fp_fn(k)
        denom = tp + 0.5 * fp + 0.5 * fn
        return tp / denom if denom else 0.0


# ---------------------------------------------------------------------------------------------------------------------
# device-resident post-processing
# ---------------------------------------------------------------------------------------------------------------------
def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('tdeed_b200 util.eval needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def _dicts(frames, labels, scores, name_of):
    return [{'label': name_of[int(l)], 'frame': int(f), 'score': float(s)} for f, l, s in zip(frames, labels, scores)]


class _VideoEvents:
    """Events of one video, still on the device (output of tdeed_extract_events)."""

    def __init__(self, video, fps, vs, threshold):
        self.video, self.fps, self.vs, self.k = video, fps, vs, vs.k
        self.ev = vs.events(threshold)
        self._counts = torch.empty(2, dtype=torch.int32, pin_memory=True)
        self._counts.copy_(self.ev['counts'], non_blocking=True)
        self._done = torch.cuda.Event()
        self._done.record()

    def lists(self, classes_inv):
        """(events, events_high_recall) in the reference's list-of-dicts form."""
        self._done.synchronize()
        out = []
        for tag, n in zip(('ev', 'hr'), self._counts.tolist()):
            out.append(_dicts(self.ev[tag + '_frame'][:n].cpu().numpy(), self.ev[tag + '_label'][:n].cpu().numpy(),
                              self.ev[tag + '_score'][:n].cpu().numpy(), classes_inv))
        return out

    def nms(self, window, threshold, soft):
        """(soft-)NMS of the high-recall events straight from the device buffers; D2H is asynchronous."""
        return PendingEvents(self.ev, self.k, window, threshold, soft)


def _frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold, with_labels):
    classes_inv = {v: k for k, v in classes.items()}
    fps_dict = {video: fps for video, _, fps in dataset.videos}
    err, f1 = ErrorStat(), ForegroundF1()
    pred_events, pred_events_high_recall, pred_scores = [], [], {}
    dev = _device()
    for video, (scores, support) in sorted(pred_dict.items()):
        if isinstance(scores, VideoScores):
            vs = scores
        else:                       # numpy buffers as in the reference: upload, normalise/extract on the device
            vs = VideoScores(scores.shape[0], scores.shape[1], dev)
            vs.scores.copy_(torch.as_tensor(scores))
            vs.support.copy_(torch.as_tensor(support))
        ve = _VideoEvents(video, fps_dict[video], vs, high_recall_score_threshold)
        events, events_high_recall = ve.lists(classes_inv)
        host_scores = vs.scores.cpu().numpy()
        if not isinstance(scores, VideoScores):      # the reference normalises its buffers in place
            scores[...] = host_scores
            support[...] = vs.support.cpu().numpy()
        pred_scores[video] = host_scores.tolist()
        if with_labels:
            label = dataset.get_labels(video)
            pred = ve.ev['pred'].cpu().numpy()
            err.update(label, pred)
            f1.update(label, pred)
        pred_events.append({'video': video, 'events': events, 'fps': fps_dict[video]})
        pred_events_high_recall.append({'video': video, 'events': events_high_recall, 'fps': fps_dict[video]})
    return err, f1, pred_events, pred_events_high_recall, pred_scores


def process_frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold=0.01):
    return _frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold, True)


def process_frame_predictions_challenge(dataset, classes, pred_dict, high_recall_score_threshold=0.05):
    return _frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold, False)[2:]


def _nms_dicts(pred, window, threshold, soft):
    """Shared driver of the two NMS flavours on the reference's list-of-dicts format: all videos are launched before the
    first result is read back (one host sync in total instead of one per video)."""
    if isinstance(window, list):
        raise NotImplementedError('per-label window lists are unused by the reference callers and not built')
    dev = _device()
    launched = []
    for video_pred in pred:
        ev = video_pred['events']
        labels = list(dict.fromkeys(e['label'] for e in ev))          # dense ids in first-appearance order
        ids = {l: i + 1 for i, l in enumerate(labels)}
        res = None
        if ev:
            sc64 = np.asarray([e['score'] for e in ev], np.float64)
            sc32 = sc64.astype(np.float32)
            if not np.array_equal(sc32.astype(np.float64), sc64):
                raise ValueError('scores must be float32-representable (they are in the reference pipeline)')
            dev_ev = {
                'hr_frame': torch.as_tensor(np.asarray([e['frame'] for e in ev], np.int32)).to(dev, non_blocking=True),
                'hr_label': torch.as_tensor(np.asarray([ids[e['label']] for e in ev], np.int32)).to(dev, non_blocking=True),
                'hr_score': torch.as_tensor(sc32).to(dev, non_blocking=True),
                'counts': torch.tensor([0, len(ev)], dtype=torch.int32).to(dev, non_blocking=True),
            }
            res = PendingEvents(dev_ev, max(len(labels) + 1, 2), window, threshold, soft)
        launched.append((video_pred, labels, res))
    new_pred = []
    for video_pred, labels, res in launched:
        new_video_pred = {key: val for key, val in video_pred.items() if key != 'events'}
        events = _dicts(*res.get(), {i + 1: l for i, l in enumerate(labels)}) if res is not None else []
        new_video_pred['events'] = events
        new_video_pred['num_events'] = len(events)
        new_pred.append(new_video_pred)
    return new_pred


def non_maximum_supression(pred, window, threshold=0.0):
    return _nms_dicts(pred, window, threshold, soft=False)


def soft_non_maximum_supression(pred, window, threshold=0.01):
    return _nms_dicts(pred, window, threshold, soft=True)


# ---------------------------------------------------------------------------------------------------------------------
# clip scores: video-level frame stream (fast path) or the reference's clip loop (any model with .predict)
# ---------------------------------------------------------------------------------------------------------------------
class _FramePieces(Dataset):
    """Item k = frames [k*P, (k+1)*P) of the videos' frames laid back to back (every video contributes exactly video_len
    frames, sub-sampled by the dataset stride) as uint8 (n,3,H,W) — each unique frame is decoded ONCE, where the reference's
    clip dataset decodes it once per overlapping clip (dataset/frame.py:425-452).  Frames whose file is missing are black,
    like the zero padding of dataset/frame.py:600-625."""

    def __init__(self, dataset, videos, piece):
        self.reader, self.stride = dataset._frame_reader, dataset._stride
        self.videos = videos                     # [(name, video_len, first clip tuple)]
        self.base = np.concatenate([[0], np.cumsum([v[1] for v in videos])]).astype(np.int64)
        self.P = piece
        self._hw = None

    def __len__(self):
        return int(math.ceil(self.base[-1] / self.P))

    def _read(self, name, clip0, j0, j1):
        kw = {'source_info': clip0[2]} if len(clip0) > 2 else {}
        fr = self.reader.load_frames(name, j0 * self.stride, j1 * self.stride, pad=False, stride=self.stride, **kw)
        return None if isinstance(fr, int) else fr

    def _load(self, name, clip0, j0, j1):
        n = j1 - j0
        fr = self._read(name, clip0, j0, j1)
        if fr is not None:
            self._hw = tuple(fr.shape[1:])
            if fr.shape[0] == n:
                return fr
        # some files are missing: read frame by frame so that every frame keeps its own index
        if self._hw is None:
            probe = self._read(name, clip0, 0, 1)
            if probe is None:
                raise RuntimeError('cannot determine the frame size of video %r (no readable frame)' % (name,))
            self._hw = tuple(probe.shape[1:])
        out = torch.zeros((n,) + self._hw, dtype=torch.uint8)
        if fr is not None:
            for i in range(n):
                one = self._read(name, clip0, j0 + i, j0 + i + 1)
                if one is not None:
                    out[i] = one[0]
        return out

    def __getitem__(self, k):
        g0, g1 = k * self.P, min(int(self.base[-1]), (k + 1) * self.P)
        parts = []
        v = int(np.searchsorted(self.base, g0, side='right')) - 1
        while g0 < g1:
            name, vlen, clip0 = self.videos[v]
            j0 = g0 - int(self.base[v])
            j1 = min(vlen, j0 + (g1 - g0))
            if j1 > j0:
                parts.append(self._load(name, clip0, j0, j1))
                g0 += j1 - j0
            v += 1
        return parts[0] if len(parts) == 1 else torch.cat(parts)


def _is_native(model):
    impl = getattr(model, '_model', None)
    return impl is not None and hasattr(impl, 'engine') and hasattr(impl, 'engine_config')


def _streamable(dataset):
    return all(hasattr(dataset, a) for a in ('_frame_reader', '_clips', '_stride', '_clip_len', 'videos'))


def _stream_scores(model, dataset, mine, augment, k, dev):
    """Fast path: {video: VideoScores} through tdeed_b200.pipeline.VideoInference (each unique frame decoded, uploaded and
    run through stem + s1 + s2 once; clips assembled from the feature ring)."""
    impl = model._model
    impl.eval()
    eng = impl.engine('bf16')                    # model.predict(use_amp=True) of the reference's loop
    stride, T = dataset._stride, dataset._clip_len
    per_video = {}
    for c in dataset._clips:
        if mine is None or c[0] in mine:
            per_video.setdefault(c[0], []).append(c)
    vlen = {name: n for name, n, _ in dataset.videos}
    order = list(per_video)
    vlist = [(name, vlen[name], [c[1] // stride for c in per_video[name]]) for name in order]
    starts = vlist[0][2] if vlist else []
    hop = (starts[1] - starts[0]) if len(starts) > 1 else max(1, T // 4)
    pieces = _FramePieces(dataset, [(name, vlen[name], per_video[name][0]) for name in order], STREAM_PIECE_FRAMES)
    if len(pieces) == 0:
        return {}
    first = pieces[0]
    in_hw = tuple(first.shape[-2:])
    ch, cw = eng.crop_window(*in_hw)[2:]
    B = max(2, min(STREAM_CLIPS_PER_BATCH, int(STREAM_CLIPS_PER_BATCH * (224 * 224) / (ch * cw))))
    key = (id(eng), in_hw, B, hop, bool(augment))
    vi = impl.__dict__.setdefault('_video_inference', {}).get(key)
    if vi is None or vi.eng is not eng:
        impl.__dict__['_video_inference'] = {key: VideoInference(eng, in_hw, clips_per_batch=B, frames_per_chunk=max(T, B * hop),
                                                                 flips=(False, True) if augment else (False,), clip_len=T)}
        vi = impl.__dict__['_video_inference'][key]
    # frames are decoded by worker threads straight into pinned buffers (tdeed_b200.pipeline.ThreadedFrameSource)
    source = ThreadedFrameSource(pieces, vi.stream, workers=STREAM_WORKERS, slots=vi.__dict__.setdefault('_host_slots', []))
    with torch.no_grad():
        return vi.run(vlist, tqdm(source, total=len(pieces)))


def _clip_loop_scores(model, dataset, videos, augment, k, dev):
    """The reference's clip loop (util/eval.py:289-349) for any model with .predict(); only the accumulation moved to the
    device (bit-exact fp32 adds in clip order)."""
    pred = {video: VideoScores(video_len, k, dev) for video, video_len, _ in videos}
    batch_size = 1 if augment else INFERENCE_BATCH_SIZE
    for clip in tqdm(DataLoader(dataset, num_workers=4 * 2, pin_memory=True, batch_size=batch_size)):
        starts = [int(s) for s in clip['start']]
        if batch_size > 1:
            _, probs = model.predict(clip['frame'])
            probs = torch.as_tensor(probs).to(dev)
            for video in dict.fromkeys(clip['video']):       # clips of one batch may belong to different videos
                sel = [i for i, v in enumerate(clip['video']) if v == video]
                pred[video].add(probs[sel].contiguous(), [starts[i] for i in sel], tta=False)
        else:
            for flip in (False, True):
                _, probs = model.predict(clip['frame'], augment_inference=flip) if flip else model.predict(clip['frame'])
                pred[clip['video'][0]].add(torch.as_tensor(probs).to(dev), starts * probs.shape[0], tta=True)
    return pred


def _windows_for(name):
    tolerances, windows = TOLERANCES, WINDOWS
    if name == 'soccernet':
        tolerances, windows = TOLERANCES_SN, WINDOWS_SN
    if name == 'soccernetball':
        tolerances, windows = TOLERANCES_SNB, WINDOWS_SNB
    if name == 'tennis':
        windows = WINDOWS_T
    if name == 'finegym':
        windows = WINDOWS_FG
    return tolerances, windows


def evaluate(model, dataset, split, classes, save_pred=None, printed=True, test=False, augment=False):
    """Same contract as util/eval.py:264-419 of the reference.  Differences are mechanical: with this repo's TDEEDModel and
    the reference's ActionSpotVideoDataset the clips come from the video-level engine (identical scores, every unique
    frame processed once); predictions stay on the device (tdeed_clip_accumulate -> tdeed_extract_events -> tdeed_nms);
    under torchrun every rank owns whole videos (bit-identical events, results merged on all ranks, files written by
    rank 0).  mAP scoring and the JSON writers are `util.score` / `util.io`."""
    tolerances, windows = _windows_for(dataset._dataset)
    dev = _device()
    k = len(classes) + 1
    classes_inv = {v: kk for kk, v in classes.items()}
    rank, ws = world()
    videos = list(dataset.videos)
    mine = None
    all_clips = None
    if ws > 1:
        counts = {}
        for c in dataset._clips:
            counts[c[0]] = counts.get(c[0], 0) + 1
        mine = shard_videos(counts.items(), rank, ws)
        videos = [v for v in videos if v[0] in mine]
    if os.environ.get('TDEED_EVAL_TIMING') == '1':
        import time
        torch.cuda.synchronize()
        _t0 = time.perf_counter()
    try:
        if _is_native(model) and _streamable(dataset):
            scores = _stream_scores(model, dataset, mine, augment, k, dev)
        else:
            if ws > 1:          # the clip loop iterates the dataset itself: restrict it to this rank's videos meanwhile
                all_clips, dataset._clips = dataset._clips, [c for c in dataset._clips if c[0] in mine]
            scores = _clip_loop_scores(model, dataset, videos, augment, k, dev)
    finally:
        if all_clips is not None:
            dataset._clips = all_clips
    if os.environ.get('TDEED_EVAL_TIMING') == '1':
        torch.cuda.synchronize()
        print('[util.eval] %-28s %.1f ms' % ('clip scores (network)', (time.perf_counter() - _t0) * 1e3), file=sys.stderr)

    timing = os.environ.get('TDEED_EVAL_TIMING') == '1'
    if timing:
        import time
        torch.cuda.synchronize()
        _t = [time.perf_counter()]

        def lap(what):
            torch.cuda.synchronize()
            _t.append(time.perf_counter())
            print('[util.eval] %-28s %.1f ms' % (what, (_t[-1] - _t[-2]) * 1e3), file=sys.stderr)
    else:
        def lap(what):
            pass
    fps = {video: f for video, _, f in videos}
    results = [_VideoEvents(video, fps[video], scores[video], 0.01) for video in sorted(scores)]
    lap('event extraction')

    def gathered(lists):
        return gather_video_results(lists) if ws > 1 else lists

    def nms_lists(window, threshold, soft):
        pend = [(r, r.nms(window, threshold, soft)) for r in results]
        out = []
        for r, p in pend:
            ev = _dicts(*p.get(), classes_inv)
            out.append({'video': r.video, 'events': ev, 'fps': r.fps, 'num_events': len(ev)})
        return gathered(out)

    from util.score import compute_mAPs
    if not test:
        nms = nms_lists(windows[0], 0.10, False)
        lap('NMS + event lists')
        mAPs, _ = compute_mAPs(dataset.labels, nms, tolerances=tolerances, printed=True)
        lap('compute_mAPs')
        return np.mean(mAPs)

    from util.io import store_json, store_json_snb, store_json_sn
    high_recall = gathered([{'video': r.video, 'events': r.lists(classes_inv)[1], 'fps': r.fps} for r in results])
    if split != 'CHALLENGE':
        err, f1 = ErrorStat(), ForegroundF1()
        for r in results:
            label = dataset.get_labels(r.video)
            pred = r.ev['pred'].cpu().numpy()
            err.update(label, pred)
            f1.update(label, pred)
        if ws > 1:                                        # frame-level statistics are sums: merge them too
            stats = [None] * ws
            torch.distributed.all_gather_object(stats, (err._err, err._total, f1._c))
            err, f1 = ErrorStat(), ForegroundF1()
            for e_, t_, c_ in stats:
                err._err += e_
                err._total += t_
                f1.merge_counts(c_)

        print('=== Results on {} (w/o NMS) ==='.format(split))
        print('Error (frame-level): {:0.2f}\n'.format(err.get() * 100))
        from tabulate import tabulate
        rows = [['any', f1.get(None) * 100, *f1.tp_fp_fn(None)]]
        for c in sorted(classes):
            rows.append([c, f1.get(classes[c]) * 100, *f1.tp_fp_fn(classes[c])])
        print(tabulate(rows, headers=['Exact frame', 'F1', 'TP', 'FP', 'FN'], floatfmt='0.2f'))
        print()

        mAPs, _ = compute_mAPs(dataset.labels, high_recall, tolerances=tolerances, printed=printed)
        print('=== Results on {} (w/ NMS{}) ==='.format(split, str(windows[0])))
        nms = nms_lists(windows[0], 0.01, False)
        mAPs, tolerances = compute_mAPs(dataset.labels, nms, tolerances=tolerances, printed=printed)
        avg_mAP_nms = np.mean(mAPs)
        print('=== Results on {} (w/ SNMS{}) ==='.format(split, str(windows[1])))
        snms = nms_lists(windows[1], 0.01, True)
        mAPs, _ = compute_mAPs(dataset.labels, snms, tolerances=tolerances, printed=printed)
        store = snms if np.mean(mAPs) > avg_mAP_nms else nms
        print('Storing predictions with SNMS' if store is snms else 'Storing predictions with NMS')
        if save_pred is not None:
            if rank == 0:
                os.makedirs(os.path.dirname(save_pred) or '.', exist_ok=True)
                store_json(save_pred + '.json', store)
                if dataset._dataset == 'soccernet':
                    store_json_sn(save_pred, store, stride=dataset._stride)
                if dataset._dataset == 'soccernetball':
                    store_json_snb(save_pred, store, stride=dataset._stride)
            if ws > 1:
                torch.distributed.barrier()
        return mAPs, tolerances

    nms_lists(windows[1], 0.01, True)      # computed and dropped, as in the reference (util/eval.py:414-418 stores the un-suppressed list)
    print('Storing predictions Challenge with SNMS')
    if rank == 0:
        store_json_snb(save_pred, high_recall, stride=dataset._stride)
    if ws > 1:
        torch.distributed.barrier()
    return None, None
