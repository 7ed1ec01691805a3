"""Drop-in for the hot-path half of the reference's util/eval.py (`evaluate`, `process_frame_predictions[_challenge]`,
`non_maximum_supression`, `soft_non_maximum_supression`, the window/tolerance constants) with the same signatures
and list-of-dict wire format, running on the device-resident kernels of libtdeed_sm100 (tdeed_clip_accumulate,
tdeed_extract_events, tdeed_nms).

`t-deed_b200/util/` deliberately has NO __init__.py: like the reference's `util/` it is a namespace-package portion,
so with `t-deed_b200` first on sys.path `util.eval` resolves here while `util.io`, `util.score`, `util.dataset`
(scoring / JSON writers — out of the hot path) keep resolving to the reference checkout.
"""
from collections import defaultdict

import numpy as np
import torch
from torch.utils.data import DataLoader
from tqdm import tqdm

from tdeed_b200 import ops
from tdeed_b200.parallel import gather_video_results, shard_videos, world
from tdeed_b200.pipeline import VideoScores

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


class ErrorStat:

    def __init__(self):
        self._total = 0
        self._err = 0

    def update(self, true, pred):
        self._err += np.sum(true != pred)
        self._total += true.shape[0]

    def get(self):
        return self._err / self._total

    def get_acc(self):
        return 1. - self.get()


class ForegroundF1:

    def __init__(self):
        self._tp = defaultdict(int)
        self._fp = defaultdict(int)
        self._fn = defaultdict(int)

    def update(self, true, pred):
        if pred != 0:
            if true != 0:
                self._tp[None] += 1
            else:
                self._fp[None] += 1
            if pred == true:
                self._tp[pred] += 1
            else:
                self._fp[pred] += 1
                if true != 0:
                    self._fn[true] += 1
        elif true != 0:
            self._fn[None] += 1
            self._fn[true] += 1

    def get(self, k):
        return self._f1(k)

    def tp_fp_fn(self, k):
        return self._tp[k], self._fp[k], self._fn[k]

    def _f1(self, k):
        denom = self._tp[k] + 0.5 * self._fp[k] + 0.5 * self._fn[k]
        if denom == 0:
            assert self._tp[k] == 0
            denom = 1
        return self._tp[k] / denom


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('tdeed_b200 util.eval needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def _events_from_device(ev, classes_inv):
    """Device event buffers of tdeed_extract_events -> the reference's two event lists (host dicts)."""
    n_ev, n_hr = ev['counts'].cpu().tolist()
    out = []
    for tag, n in (('ev', n_ev), ('hr', n_hr)):
        fr = ev[tag + '_frame'][:n].cpu().numpy()
        lb = ev[tag + '_label'][:n].cpu().numpy()
        sc = ev[tag + '_score'][:n].cpu().numpy()
        out.append([{'label': classes_inv[int(l)], 'frame': int(f), 'score': float(s)} for f, l, s in zip(fr, lb, sc)])
    return out


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
        ev = vs.events(high_recall_score_threshold)
        events, events_high_recall = _events_from_device(ev, classes_inv)
        host_scores = vs.scores.cpu().numpy()
        if not isinstance(scores, VideoScores):      # the reference normalises its buffers in place
            scores[...] = host_scores
            support[...] = vs.support.cpu().numpy()
        pred_scores[video] = host_scores.tolist()
        if with_labels:
            label = dataset.get_labels(video)
            pred = ev['pred'].cpu().numpy()
            err.update(label, pred)
            for i in range(pred.shape[0]):
                f1.update(label[i], pred[i])
        pred_events.append({'video': video, 'events': events, 'fps': fps_dict[video]})
        pred_events_high_recall.append({'video': video, 'events': events_high_recall, 'fps': fps_dict[video]})
    return err, f1, pred_events, pred_events_high_recall, pred_scores


def process_frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold=0.01):
    return _frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold, True)


def process_frame_predictions_challenge(dataset, classes, pred_dict, high_recall_score_threshold=0.05):
    return _frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold, False)[2:]


def _nms_dicts(pred, window, threshold, soft):
    """Shared driver of the two NMS flavours on the reference's list-of-dicts format."""
    if isinstance(window, list):
        raise NotImplementedError('per-label window lists are unused by the reference callers and not built')
    dev = _device()
    new_pred = []
    for video_pred in pred:
        ev = video_pred['events']
        labels = []
        for e in ev:                                   # label names -> dense ids in first-appearance order
            if e['label'] not in labels:
                labels.append(e['label'])
        ids = {l: i + 1 for i, l in enumerate(labels)}
        k = len(labels) + 1
        new_video_pred = {key: val for key, val in video_pred.items() if key != 'events'}
        if ev:
            fr = torch.as_tensor(np.asarray([e['frame'] for e in ev], np.int32)).to(dev)
            lb = torch.as_tensor(np.asarray([ids[e['label']] for e in ev], np.int32)).to(dev)
            sc64 = np.asarray([e['score'] for e in ev], np.float64)
            sc32 = sc64.astype(np.float32)
            if not np.array_equal(sc32.astype(np.float64), sc64):
                raise ValueError('scores must be float32-representable (they are in the reference pipeline)')
            cnt = torch.tensor([len(ev)], dtype=torch.int32, device=dev)
            of, ol, os_, oc = ops.nms(fr, lb, torch.as_tensor(sc32).to(dev), cnt, max(k, 2), window, threshold, soft)
            n = int(oc.item())
            events = [{'label': labels[int(l) - 1], 'frame': int(f), 'score': float(s)}
                      for f, l, s in zip(of[:n].cpu().numpy(), ol[:n].cpu().numpy(), os_[:n].cpu().numpy())]
        else:
            events = []
        new_video_pred['events'] = events
        new_video_pred['num_events'] = len(events)
        new_pred.append(new_video_pred)
    return new_pred


def non_maximum_supression(pred, window, threshold=0.0):
    return _nms_dicts(pred, window, threshold, soft=False)


def soft_non_maximum_supression(pred, window, threshold=0.01):
    return _nms_dicts(pred, window, threshold, soft=True)


def evaluate(model, dataset, split, classes, save_pred=None, printed=True, test=False, augment=False):
    """Same contract as util/eval.py:264-419 of the reference.  Differences are mechanical: predictions stay on
    the device (engine -> tdeed_clip_accumulate -> tdeed_extract_events -> tdeed_nms); mAP scoring and the JSON
    writers are the reference's own (imported lazily from util.score / util.io)."""
    tolerances, windows = TOLERANCES, WINDOWS
    if dataset._dataset == 'soccernet':
        tolerances, windows = TOLERANCES_SN, WINDOWS_SN
    if dataset._dataset == 'soccernetball':
        tolerances, windows = TOLERANCES_SNB, WINDOWS_SNB
    if dataset._dataset == 'tennis':
        windows = WINDOWS_T
    if dataset._dataset == 'finegym':
        windows = WINDOWS_FG

    dev = _device()
    k = len(classes) + 1
    # multi-GPU (torchrun): every rank owns whole videos -> accumulation order and events are bit-identical to 1 GPU
    rank, ws = world()
    all_clips = None
    videos = list(dataset.videos)
    if ws > 1:
        counts = defaultdict(int)
        for c in dataset._clips:
            counts[c[0]] += 1
        mine = shard_videos(counts.items(), rank, ws)
        all_clips, dataset._clips = dataset._clips, [c for c in dataset._clips if c[0] in mine]
        videos = [v for v in videos if v[0] in mine]
    pred_dict = {video: (VideoScores(video_len, k, dev), None) for video, video_len, _ in videos}
    batch_size = 1 if augment else INFERENCE_BATCH_SIZE
    impl = model._model
    impl.eval()

    def probs_on_device(frames, flip):
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            impl(frames.to(dev, non_blocking=True), inference=True, augment_inference=flip, use_graph=True)
        return impl._last_probs

    for clip in tqdm(DataLoader(dataset, num_workers=4 * 2, pin_memory=True, batch_size=batch_size)):
        starts = [int(s) for s in clip['start']]
        if batch_size > 1:
            probs = probs_on_device(clip['frame'], False)
            # clips of one batch may belong to different videos: accumulate per video, in clip order
            for video in dict.fromkeys(clip['video']):
                sel = [i for i, v in enumerate(clip['video']) if v == video]
                pred_dict[video][0].add(probs[sel].contiguous(), [starts[i] for i in sel], tta=False)
        else:
            vs = pred_dict[clip['video'][0]][0]
            for flip in (False, True):
                vs.add(probs_on_device(clip['frame'], flip).clone(), starts, tta=True)

    if split != 'CHALLENGE':
        err, f1, pred_events, pred_events_high_recall, pred_scores = \
            process_frame_predictions(dataset, classes, pred_dict, high_recall_score_threshold=0.01)
    else:
        pred_events, pred_events_high_recall, pred_scores = \
            process_frame_predictions_challenge(dataset, classes, pred_dict, high_recall_score_threshold=0.01)
    if ws > 1:
        dataset._clips = all_clips
        pred_events = gather_video_results(pred_events)
        pred_events_high_recall = gather_video_results(pred_events_high_recall)
        if split != 'CHALLENGE':                      # frame-level error statistics are sums: merge them too
            stats = [None] * ws
            torch.distributed.all_gather_object(stats, (err._err, err._total, dict(f1._tp), dict(f1._fp), dict(f1._fn)))
            err, f1 = ErrorStat(), ForegroundF1()
            for e_, t_, tp, fp, fn in stats:
                err._err += e_
                err._total += t_
                for src, dst in ((tp, f1._tp), (fp, f1._fp), (fn, f1._fn)):
                    for key, val in src.items():
                        dst[key] += val

    from util.score import compute_mAPs            # the reference's scorer (not on the device path)
    if not test:
        pred_events_high_recall = non_maximum_supression(pred_events_high_recall, window=windows[0], threshold=0.10)
        mAPs, _ = compute_mAPs(dataset.labels, pred_events_high_recall, tolerances=tolerances, printed=True)
        return np.mean(mAPs)

    from util.io import store_json, store_json_sn, store_json_snb
    import os
    if split != 'CHALLENGE':
        print('=== Results on {} (w/o NMS) ==='.format(split))
        print('Error (frame-level): {:0.2f}\n'.format(err.get() * 100))
        mAPs, _ = compute_mAPs(dataset.labels, pred_events_high_recall, tolerances=tolerances, printed=printed)
        print('=== Results on {} (w/ NMS{}) ==='.format(split, str(windows[0])))
        nms = non_maximum_supression(pred_events_high_recall, window=windows[0], threshold=0.01)
        mAPs, tolerances = compute_mAPs(dataset.labels, nms, tolerances=tolerances, printed=printed)
        avg_mAP_nms = np.mean(mAPs)
        print('=== Results on {} (w/ SNMS{}) ==='.format(split, str(windows[1])))
        snms = soft_non_maximum_supression(pred_events_high_recall, window=windows[1], threshold=0.01)
        mAPs, _ = compute_mAPs(dataset.labels, snms, tolerances=tolerances, printed=printed)
        store = snms if np.mean(mAPs) > avg_mAP_nms else nms
        print('Storing predictions with SNMS' if store is snms else 'Storing predictions with NMS')
        if save_pred is not None:
            os.makedirs(os.path.dirname(save_pred) or '.', exist_ok=True)
            store_json(save_pred + '.json', store)
            if dataset._dataset == 'soccernet':
                store_json_sn(save_pred, store, stride=dataset._stride)
            if dataset._dataset == 'soccernetball':
                store_json_snb(save_pred, store, stride=dataset._stride)
        return mAPs, tolerances

    soft_non_maximum_supression(pred_events_high_recall, window=windows[1], threshold=0.01)
    print('Storing predictions Challenge with SNMS')
    store_json_snb(save_pred, pred_events_high_recall, stride=dataset._stride)   # reference quirk (:416-418): un-suppressed list
    return None, None
