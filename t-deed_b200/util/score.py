"""Drop-in for the reference's util/score.py (mAP scoring): `compute_mAPs` with the same signature,
return value and printed table, computed as

  host   label / video bucketing, stable descending-score order (== list.sort(key=score, reverse=True)), the
         precision curve, its right-to-left running maximum and the left-to-right float64 sum (the reference's exact
         summation order, so every AP is the same double);
  device the greedy prediction <-> ground-truth matching for ALL (class, video, tolerance) triples in one launch of
         tdeed_match_events (csrc/score.cu) — the O(P x G) pure-Python double loop of util/score.py:57-75, which is
         sequential only inside one (class, video) pair.

No CPU fallback for the matching: without the CUDA library compute_mAPs raises.
"""
import os
import sys
from collections import defaultdict

import numpy as np
import torch
from tabulate import tabulate

from tdeed_b200 import ops
from util.io import load_json, load_text

FPS_SN = 25


def parse_ground_truth(truth):
    """{label: {video: [frames in label-file order]}} (util/score.py:16-32); SoccerNet entries without 'events' read the
    game's Labels-v2.json, events without 'frame' convert `position` (ms) at FPS_SN."""
    label_dict = defaultdict(lambda: defaultdict(list))
    for x in truth:
        if 'events' in x:
            events = x['events']
        else:
            root = load_text(os.path.join('data', 'soccernet', 'labels_path.txt'))[0]
            events = load_json(os.path.join(root, '/'.join(x['video'].split('/')[:-1]) + '/Labels-v2.json'))['annotations']
        for e in events:
            frame = e['frame'] if 'frame' in e else int(int(e['position']) / 1000 * FPS_SN)
            label_dict[e['label']][x['video']].append(frame)
    return label_dict


def get_predictions(pred, label=None):
    """[(video, frame, score)] of one label (all if None), best score first, ties in list order (util/score.py:35-42)."""
    flat = [(x['video'], e['frame'], e['score']) for x in pred for e in x['events'] if label is None or e['label'] == label]
    flat.sort(key=lambda t: t[-1], reverse=True)
    return flat


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('tdeed_b200 util.score needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def compute_average_precisions(pred, truth_by_label, labels, tolerances):
    """{(label, tolerance): AP} for all labels x tolerances (util/score.py:45-89 for each pair)."""
    dev = _device()
    vid_index = {x['video']: i for i, x in enumerate(pred)}
    inv_video = [x['video'] for x in pred]
    flat_frames, unit_lo, unit_hi, gt_frames, gt_off = [], [], [], [], [0]
    layout, total = {}, 0
    for l in labels:
        vids, frames, scores = [], [], []
        for x in pred:
            vi = vid_index[x['video']]
            for e in x['events']:
                if e['label'] == l:
                    vids.append(vi)
                    frames.append(e['frame'])
                    scores.append(e['score'])
        vids = np.asarray(vids, np.int64)
        frames = np.asarray(frames, np.int64)
        scores = np.asarray(scores, np.float64)
        assert scores.size == 0 or scores.max() <= 1, 'scores must be <= 1 (util/score.py:55)'
        order = np.argsort(-scores, kind='stable')                 # == sort(key=score, reverse=True): ties keep list order
        by_video = np.argsort(vids[order], kind='stable')          # one segment per video, score order kept inside
        seg_vid = vids[order][by_video]
        flat_frames.append(frames[order][by_video])
        if seg_vid.size:
            cuts = np.flatnonzero(np.diff(seg_vid)) + 1
            for s, e in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [seg_vid.size]])):
                gts = truth_by_label[l].get(inv_video[int(seg_vid[s])], [])
                if gts:                                            # videos without ground truth: every prediction is a false positive
                    unit_lo.append(total + int(s))
                    unit_hi.append(total + int(e))
                    gt_frames.extend(int(v) for v in gts)
                    gt_off.append(len(gt_frames))
        layout[l] = (total, frames.size, order, by_video)
        total += frames.size
    n_tol = len(tolerances)
    if unit_lo and total:
        # the kernel takes CSR offsets: make the units contiguous by gathering their predictions
        sel = np.concatenate([np.arange(a, b) for a, b in zip(unit_lo, unit_hi)])
        allf = np.concatenate(flat_frames)
        poff = np.concatenate([[0], np.cumsum(np.asarray(unit_hi) - np.asarray(unit_lo))])
        i32 = lambda a: torch.as_tensor(np.asarray(a, np.int32)).to(dev, non_blocking=True)
        tp_units = ops.match_events(i32(allf[sel]), i32(poff), i32(gt_frames), i32(gt_off), i32(list(tolerances))).cpu().numpy()
        tp = np.zeros((n_tol, total), np.uint8)
        tp[:, sel] = tp_units
    else:
        tp = np.zeros((n_tol, total), np.uint8)
    aps = {}
    for l in labels:
        base, n, order, by_video = layout[l]
        n_truth = sum(len(v) for v in truth_by_label[l].values())
        for ti, tol in enumerate(tolerances):
            hit = np.zeros(n, bool)
            hit[by_video] = tp[ti, base:base + n] != 0            # back to descending-score order
            ranks = np.flatnonzero(hit) + 1                       # 1-based position of every true positive
            pc = np.arange(1, ranks.size + 1) / ranks             # precision each time recall grows (float64, as len/i)
            interp = np.maximum.accumulate(pc[::-1])[::-1]
            aps[(l, tol)] = sum(interp.tolist()) / n_truth        # left-to-right double sum like the reference
    return aps


def compute_average_precision(pred, truth, tolerance=0, min_precision=0, plot_ax=None, plot_label=None, plot_raw_pr=True):
    """Single (label, tolerance) entry point with the reference's signature (util/score.py:45-89): pred = the
    get_predictions(...) list (already in descending-score order), truth = {video: [frames]}."""
    if min_precision:
        raise NotImplementedError('min_precision early stop is unused by the reference callers')
    return _average_precision_sorted(pred, truth, tolerance, plot_ax, plot_label, plot_raw_pr)


def _average_precision_sorted(pred_sorted, truth, tolerance, plot_ax=None, plot_label=None, plot_raw_pr=True):
    dev = _device()
    n = len(pred_sorted)
    vids = list(dict.fromkeys(v for v, _, _ in pred_sorted))
    vidx = {v: i for i, v in enumerate(vids)}
    vi = np.asarray([vidx[v] for v, _, _ in pred_sorted], np.int64)
    fr = np.asarray([f for _, f, _ in pred_sorted], np.int64)
    by_video = np.argsort(vi, kind='stable')
    seg = vi[by_video]
    hit = np.zeros(n, bool)
    if n:
        cuts = np.flatnonzero(np.diff(seg)) + 1
        lo_hi = [(int(s), int(e)) for s, e in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [n]]))
                 if truth.get(vids[int(seg[s])])]
        if lo_hi:
            sel = np.concatenate([np.arange(a, b) for a, b in lo_hi])
            poff = np.concatenate([[0], np.cumsum([b - a for a, b in lo_hi])])
            gts, goff = [], [0]
            for a, _ in lo_hi:
                gts.extend(int(x) for x in truth[vids[int(seg[a])]])
                goff.append(len(gts))
            i32 = lambda a: torch.as_tensor(np.asarray(a, np.int32)).to(dev)
            tp = ops.match_events(i32(fr[by_video][sel]), i32(poff), i32(gts), i32(goff), i32([tolerance])).cpu().numpy()[0]
            flags = np.zeros(n, bool)
            flags[sel] = tp != 0
            hit[by_video] = flags
    total = sum(len(x) for x in truth.values())
    ranks = np.flatnonzero(hit) + 1
    pc = np.arange(1, ranks.size + 1) / ranks
    interp = np.maximum.accumulate(pc[::-1])[::-1] if pc.size else pc
    if plot_ax is not None:
        rc = np.arange(1, len(pc) + 1) / total
        if plot_raw_pr:
            plot_ax.plot(rc, pc, label=plot_label, alpha=0.8)
        plot_ax.plot(rc, interp, label=plot_label, alpha=0.8)
    return sum(interp.tolist()) / total


def compute_mAPs(truth, pred, tolerances=[0, 1, 2, 4], plot_pr=False, printed=False, stride=1):
    """Same contract as util/score.py:92-160: returns (mAPs per tolerance, tolerances); prints the per-class table."""
    assert {v['video'] for v in truth} == {v['video'] for v in pred}, 'Video set mismatch!'
    truth_by_label = parse_ground_truth(truth)
    labels = sorted(truth_by_label)
    axes = fig = plt = None
    if plot_pr:
        import matplotlib.pyplot as plt
        fig, axes = plt.subplots(len(labels), len(tolerances), sharex=True, sharey=True, figsize=(16, 16))
        aps = {(l, tol): compute_average_precision(get_predictions(pred, label=l), truth_by_label[l], tolerance=tol,
                                                   plot_ax=axes[j, i])
               for i, tol in enumerate(tolerances) for j, l in enumerate(labels)}
    else:
        aps = compute_average_precisions(pred, truth_by_label, labels, tolerances)
    mAPs = [np.mean([aps[(l, tol)] for l in labels]) for tol in tolerances]
    if printed:
        rows = [[l] + [aps[(l, tol)] * 100 for tol in tolerances] for l in labels]
        rows.append(['mAP'] + [m * 100 for m in mAPs])
        print(tabulate(rows, headers=['AP @ tol'] + tolerances, floatfmt='0.2f'))
        print('Avg mAP (across tolerances): {:0.2f}'.format(np.mean(mAPs) * 100))
    if plot_pr:
        for i, tol in enumerate(tolerances):
            for j, label in enumerate(labels):
                ax = axes[j, i]
                ax.set_xlabel('Recall')
                ax.set_xlim(0, 1)
                ax.set_ylabel('Precision')
                ax.set_ylim(0, 1.01)
                ax.set_title('{} @ tol={}'.format(label, tol))
        plt.tight_layout()
        plt.show()
        plt.close(fig)
    sys.stdout.flush()
    return mAPs, tolerances
