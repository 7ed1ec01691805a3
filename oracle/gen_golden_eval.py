#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — golden vectors for `util.eval.evaluate` and `util.score.compute_mAPs`, produced by the UNMODIFIED
reference functions (imported through oracle/ref_import.py) on the synthetic dataset / canned model of oracle/synth_data.py.

Run in the build container only:   python oracle/gen_golden_eval.py      -> tests/golden/evaluate.npz

For every case the reference's evaluate() runs end to end (DataLoader -> model.predict -> accumulation loop ->
process_frame_predictions -> NMS / SNMS -> compute_mAPs -> store_json*).  Captured: the accumulated per-video
(scores, support) just before normalisation (by wrapping the reference's process_frame_predictions*, which receives
them), the return value, and every file written.  Cases cover the TTA path (augment=True: the path the reference uses for
every dataset but the SoccerNet ones, train_tdeed.py:265), the batched path, stride 2, the validation (test=False) and the
CHALLENGE branches.
"""
import glob
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from ref_import import reference_modules  # noqa: E402
import synth_data as S  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

CLASSES = {'jump': 1, 'spin': 2, 'step': 3, 'fall': 4}

# name -> (dataset kwargs, evaluate kwargs)
CASES = {
    'tta_test': (dict(lengths={'vid_a': 83, 'vid_b': 131, 'vid_c': 17}, hw=(8, 8), clip_len=20, overlap_len=15, stride=1,
                      dataset='fs_comp', seed=3), dict(split='TEST', test=True, augment=True)),
    'tta_val': (dict(lengths={'vid_a': 64, 'vid_b': 97}, hw=(8, 8), clip_len=20, overlap_len=15, stride=1,
                     dataset='finediving', seed=4), dict(split='VAL', test=False, augment=True)),
    'batched_snb_test': (dict(lengths={'game/one': 301, 'game/two': 198}, hw=(8, 8), clip_len=20, overlap_len=15, stride=2,
                              dataset='soccernetball', seed=5), dict(split='TEST', test=True, augment=False)),
    'batched_snb_challenge': (dict(lengths={'game/one': 150}, hw=(8, 8), clip_len=20, overlap_len=15, stride=2,
                                   dataset='soccernetball', seed=6), dict(split='CHALLENGE', test=True, augment=False)),
    'batched_sn_test': (dict(lengths={'l/s/g2/1': 90, 'l/s/g2/2': 75, 'l/s/g3/1': 66, 'l/s/g3/2': 101}, hw=(8, 8), clip_len=20,
                             overlap_len=10, stride=1, dataset='soccernet', seed=8), dict(split='TEST', test=True, augment=False)),
    'batched_sn_val': (dict(lengths={'l/s/g1/1': 140, 'l/s/g1/2': 120}, hw=(8, 8), clip_len=20, overlap_len=10, stride=1,
                            dataset='soccernet', seed=7), dict(split='VAL', test=False, augment=False)),
}


def make(case):
    ds_kw, ev_kw = CASES[case]
    ds = S.SyntheticVideoDataset(CLASSES, **ds_kw)
    model = S.CannedModel(ds, len(CLASSES) + 1, seed=ds_kw['seed'])
    return ds, model, ev_kw


def run_case(ev_mod, case, out):
    ds, model, ev_kw = make(case)
    captured = {}

    def wrap(fn):
        def inner(dataset, classes, pred_dict, **kw):
            for video, (scores, support) in pred_dict.items():
                captured[video] = (scores.copy(), support.copy())
            return fn(dataset, classes, pred_dict, **kw)
        return inner
    orig = ev_mod.process_frame_predictions, ev_mod.process_frame_predictions_challenge
    ev_mod.process_frame_predictions, ev_mod.process_frame_predictions_challenge = wrap(orig[0]), wrap(orig[1])
    try:
        with tempfile.TemporaryDirectory() as tmp:
            save_pred = os.path.join(tmp, 'run', 'pred-test')
            buf = io.StringIO()
            with redirect_stdout(buf):
                ret = ev_mod.evaluate(model, ds, ev_kw['split'], CLASSES, save_pred if ev_kw['test'] else None, printed=True,
                                      test=ev_kw['test'], augment=ev_kw['augment'])
            files = {}
            for path in sorted(glob.glob(os.path.join(tmp, '**', '*.json'), recursive=True)):
                files[os.path.relpath(path, tmp)] = open(path).read()
    finally:
        ev_mod.process_frame_predictions, ev_mod.process_frame_predictions_challenge = orig
    p = case + '/'
    for video, (scores, support) in captured.items():
        out[p + 'scores_sum/' + video] = scores
        out[p + 'support/' + video] = support
    if ev_kw['test']:
        mAPs, tolerances = ret
        out[p + 'mAPs'] = np.asarray(mAPs if mAPs is not None else [], np.float64)
        out[p + 'tolerances'] = np.asarray(tolerances if tolerances is not None else [], np.int64)
    else:
        out[p + 'avg_mAP'] = np.asarray(ret, np.float64)
    out[p + 'files'] = np.frombuffer(json.dumps(files).encode(), np.uint8)
    out[p + 'stdout'] = np.frombuffer(buf.getvalue().encode(), np.uint8)
    print(case, 'videos', len(captured), 'ret', ret if not ev_kw['test'] else (ret[0], ret[1]), 'files', list(files))


def score_inputs(seed, n_videos=6, n_frames=400, labels=('a', 'b', 'c'), gt_per=9, pred_per=60):
    """Random truth / prediction lists with the awkward cases on purpose: tied scores, duplicate ground-truth frames,
    videos without ground truth for a class, classes predicted but absent from the truth."""
    rng = np.random.default_rng(seed)
    truth, pred = [], []
    for v in range(n_videos):
        name = 'v%02d' % v
        ev = []
        for l in labels:
            if rng.random() < 0.2:
                continue
            fr = rng.integers(0, n_frames, size=gt_per).tolist()
            if rng.random() < 0.5:
                fr.append(fr[0])                                   # duplicate ground-truth frame
            ev += [{'frame': int(f), 'label': l} for f in fr]
        order = rng.permutation(len(ev))
        truth.append({'video': name, 'num_frames': n_frames, 'fps': 25.0, 'events': [ev[i] for i in order]})
        pe = []
        for l in labels + ('zzz',):
            fr = rng.integers(0, n_frames, size=pred_per)
            sc = np.round(rng.random(pred_per), 2)                   # 2 decimals: plenty of ties
            pe += [{'label': l, 'frame': int(f), 'score': float(s)} for f, s in zip(fr, sc)]
        pe.sort(key=lambda e: e['frame'])
        pred.append({'video': name, 'events': pe, 'fps': 25.0})
    return truth, pred[::-1]


def run_score(score_mod, out):
    for ci, (seed, tolerances) in enumerate([(11, [0, 1, 2, 4]), (12, [6, 12]), (13, [1])]):
        truth, pred = score_inputs(seed)
        with redirect_stdout(io.StringIO()):
            mAPs, tol = score_mod.compute_mAPs(truth, pred, tolerances=tolerances, printed=True)
        by_label = score_mod.parse_ground_truth(truth)
        aps = [[score_mod.compute_average_precision(score_mod.get_predictions(pred, label=l), by_label[l], tolerance=t)
                for t in tolerances] for l in sorted(by_label)]
        p = 'score/%d/' % ci
        out[p + 'inputs'] = np.frombuffer(json.dumps({'truth': truth, 'pred': pred, 'tolerances': tolerances}).encode(), np.uint8)
        out[p + 'mAPs'] = np.asarray(mAPs, np.float64)
        out[p + 'aps'] = np.asarray(aps, np.float64)
        print('score case', ci, 'mAPs', mAPs)


def main():
    out = {}
    with reference_modules() as mods:
        for case in CASES:
            run_case(mods['util.eval'], case, out)
        import importlib
        run_score(importlib.import_module('util.score'), out)
    np.savez_compressed(os.path.join(GOLDEN, 'evaluate.npz'), **out)


if __name__ == '__main__':
    main()
