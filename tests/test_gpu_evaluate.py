"""`util.eval.evaluate` / `util.score.compute_mAPs` of the drop-in on the GPU against golden vectors from the UNMODIFIED
reference (tests/golden/evaluate.npz): return values, printed tables and every written file must be identical —
accumulation bit-exact in fp32 (incl. the TTA path, util/eval.py:319-349), events / NMS / SNMS bit-exact, mAPs equal as
doubles, JSON byte-identical.  Then the same through the native model: the video-level fast path must equal the
reference-style clip loop."""
import glob
import io
import json
import os
import tempfile
from argparse import Namespace
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

import score_oracle as SO
import synth_data as S
import tdeed_oracle as O
from gen_golden_eval import CASES, CLASSES, make, score_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda')


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'evaluate.npz'))


def _text(arr):
    return bytes(arr).decode()


def _run_evaluate(model, ds, ev_kw):
    import util.eval as E
    with tempfile.TemporaryDirectory() as tmp:
        save_pred = os.path.join(tmp, 'run', 'pred-test')
        buf = io.StringIO()
        with redirect_stdout(buf):
            ret = E.evaluate(model, ds, ev_kw['split'], CLASSES, save_pred if ev_kw['test'] else None, printed=True,
                             test=ev_kw['test'], augment=ev_kw['augment'])
        files = {os.path.relpath(p, tmp): open(p).read() for p in sorted(glob.glob(os.path.join(tmp, '**', '*.json'), recursive=True))}
    return ret, files, buf.getvalue()


@pytest.mark.parametrize('case', sorted(CASES))
def test_evaluate_equals_reference_golden(case, gold, dev):
    import util.eval as E
    ds, model, ev_kw = make(case)
    # accumulation (device, bit-exact) through the same loop evaluate() uses for foreign models
    k = len(CLASSES) + 1
    acc = E._clip_loop_scores(model, ds, list(ds.videos), ev_kw['augment'], k, dev)
    for v, vs in acc.items():
        assert np.array_equal(vs.scores.cpu().numpy(), gold['%s/scores_sum/%s' % (case, v)]), (case, v)
        assert np.array_equal(vs.support.cpu().numpy(), gold['%s/support/%s' % (case, v)]), (case, v)
    ret, files, out = _run_evaluate(model, ds, ev_kw)
    if ev_kw['test']:
        want = gold[case + '/mAPs'].tolist()
        if want:
            assert [float(m) for m in ret[0]] == want and list(ret[1]) == gold[case + '/tolerances'].tolist()
        else:
            assert ret == (None, None)
    else:
        assert float(ret) == float(gold[case + '/avg_mAP'])
    assert files == json.loads(_text(gold[case + '/files']))
    assert out == _text(gold[case + '/stdout'])


@pytest.mark.parametrize('ci', [0, 1, 2])
def test_compute_maps_equals_reference_golden(ci, gold, dev):
    import util.score as US
    inp = json.loads(_text(gold['score/%d/inputs' % ci]))
    mAPs, tol = US.compute_mAPs(inp['truth'], inp['pred'], tolerances=inp['tolerances'])
    assert [float(m) for m in mAPs] == gold['score/%d/mAPs' % ci].tolist() and tol == inp['tolerances']
    by_label = US.parse_ground_truth(inp['truth'])
    aps = [[US.compute_average_precision(US.get_predictions(inp['pred'], label=l), by_label[l], tolerance=t)
            for t in inp['tolerances']] for l in sorted(by_label)]
    assert aps == gold['score/%d/aps' % ci].tolist()


def test_compute_maps_equals_oracle_on_larger_random_input(dev):
    import util.score as US
    truth, pred = score_inputs(99, n_videos=12, n_frames=3000, labels=('a', 'b', 'c', 'd', 'e'), gt_per=40, pred_per=700)
    tolerances = [0, 1, 2, 4, 12]
    means, table = SO.mean_average_precisions(truth, pred, tolerances)
    mAPs, _ = US.compute_mAPs(truth, pred, tolerances=tolerances)
    assert [float(m) for m in mAPs] == means


def _native_model(cfg, dev, seed=5):
    from model.model import TDEEDModel
    args = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=cfg.radi_displacement,
                     feature_arch=cfg.feature_arch, clip_len=cfg.clip_len, n_layers=cfg.n_layers, sgp_ks=cfg.sgp_ks,
                     sgp_r=cfg.sgp_r, num_classes=cfg.num_classes, crop_dim=cfg.crop_dim)
    m = TDEEDModel(device=str(dev), args=args)
    m.load(O.random_state(cfg, seed))
    return m


class _PredictOnly:
    """Hides the native engine: evaluate() then takes the reference-style clip loop through model.predict."""

    def __init__(self, model):
        self._m = model

    def predict(self, *a, **k):
        return self._m.predict(*a, **k)


@pytest.mark.parametrize('augment,stride,dataset,jpeg', [(True, 1, 'fs_comp', True), (False, 2, 'soccernetball', False),
                                                         (False, 1, 'finediving', False)])
def test_native_fast_path_equals_clip_loop(dev, augment, stride, dataset, jpeg):
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=16, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=4, radi_displacement=1,
                   crop_dim=32 if dataset != 'soccernetball' else None)
    m = _native_model(cfg, dev)
    lengths = {'clipA': 70, 'clipB': 45, 'dir/clipC': 12} if dataset == 'soccernetball' else {'clipA': 70, 'clipB': 45, 'clipC': 12}
    with tempfile.TemporaryDirectory() as tmp:
        ds = S.SyntheticVideoDataset(CLASSES, lengths=lengths, hw=(32, 56), clip_len=16, overlap_len=12, stride=stride,
                                     dataset=dataset, seed=21, events_per_100=6.0)
        if jpeg:
            S.write_jpegs(ds, tmp)
        ev_kw = dict(split='TEST', test=True, augment=augment)
        fast = _run_evaluate(m, ds, ev_kw)
        slow = _run_evaluate(_PredictOnly(m), ds, ev_kw)
    assert [float(x) for x in fast[0][0]] == [float(x) for x in slow[0][0]]
    assert fast[1] == slow[1] and fast[1]
    assert fast[2] == slow[2]
    n_events = sum(len(v['events']) for v in json.loads(fast[1]['run/pred-test.json']))
    assert n_events > 0
