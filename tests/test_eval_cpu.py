"""CPU-side checks of the drop-in `util` package and of the post-processing / scoring oracles against golden vectors that
the UNMODIFIED reference produced (tests/golden/evaluate.npz, written by oracle/gen_golden_eval.py)."""
import io
import json
import os
import subprocess
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np
import pytest

import postproc_oracle as P
import score_oracle as SO
import synth_data as S
from gen_golden_eval import CASES, CLASSES, make

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'evaluate.npz'))


def _text(arr):
    return bytes(arr).decode()


@pytest.mark.parametrize('case', sorted(CASES))
def test_oracle_accumulation_matches_reference_evaluate(case, gold):
    """accumulate_tta / accumulate_batched replay the reference's evaluate() loop (util/eval.py:289-349) bit for bit."""
    ds, model, ev_kw = make(case)
    k = len(CLASSES) + 1
    acc = {v: (np.zeros((n, k), np.float32), np.zeros(n, np.int32)) for v, n, _ in ds.videos}
    for i in range(len(ds)):
        clip = ds[i]
        scores, support = acc[clip['video']]
        if ev_kw['augment']:
            for flip in (False, True):
                _, probs = model.predict(clip['frame'], augment_inference=flip)
                P.accumulate_tta(scores, support, probs, clip['start'])
        else:
            _, probs = model.predict(clip['frame'])
            P.accumulate_batched(scores, support, probs[0], clip['start'])
    for v in acc:
        assert np.array_equal(acc[v][0], gold['%s/scores_sum/%s' % (case, v)]), (case, v)
        assert np.array_equal(acc[v][1], gold['%s/support/%s' % (case, v)]), (case, v)


@pytest.mark.parametrize('ci', [0, 1, 2])
def test_score_oracle_matches_reference(ci, gold):
    inp = json.loads(_text(gold['score/%d/inputs' % ci]))
    means, table = SO.mean_average_precisions(inp['truth'], inp['pred'], inp['tolerances'])
    assert means == gold['score/%d/mAPs' % ci].tolist()
    labels = sorted(SO.parse_ground_truth(inp['truth']))
    got = [[table[(l, t)] for t in inp['tolerances']] for l in labels]
    assert got == gold['score/%d/aps' % ci].tolist()


@pytest.mark.parametrize('case', ['batched_snb_test', 'batched_snb_challenge', 'batched_sn_test'])
def test_wire_formats_are_byte_identical(case, gold):
    """util/io.py:14-68: our writers reproduce the reference's files byte for byte from the same event lists."""
    sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
    import importlib
    uio = importlib.import_module('util.io')
    assert uio.__file__.startswith(os.path.join(ROOT, 't-deed_b200'))
    files = json.loads(_text(gold[case + '/files']))
    stride = CASES[case][0]['stride']
    spotting = {k: v for k, v in files.items() if k.endswith('results_spotting.json')}
    assert spotting
    # rebuild the stored event list from the golden files themselves: UrlLocal + predictions -> events
    pred = []
    if 'run/pred-test.json' in files:
        pred = json.loads(files['run/pred-test.json'])
    else:                                       # CHALLENGE writes only the SoccerNet files: invert the position formula
        for text in spotting.values():
            g = json.loads(text)
            pred.append({'video': g['UrlLocal'], 'events': [
                {'label': p['label'], 'frame': round(p['position'] / 1000 * 25 / stride), 'score': p['confidence']}
                for p in g['predictions']]})
    with tempfile.TemporaryDirectory() as tmp:
        save_pred = os.path.join(tmp, 'run', 'pred-test')
        os.makedirs(os.path.dirname(save_pred))
        if 'run/pred-test.json' in files:
            uio.store_json(save_pred + '.json', pred)
            assert open(save_pred + '.json').read() == files['run/pred-test.json']
        (uio.store_json_sn if 'sn_' in case and 'snb' not in case else uio.store_json_snb)(save_pred, pred, stride=stride)
        for rel, text in spotting.items():
            assert open(os.path.join(tmp, rel)).read() == text, rel
        assert uio.load_json(save_pred + '.json') == pred if 'run/pred-test.json' in files else True
    with tempfile.NamedTemporaryFile('w', suffix='.txt', delete=False) as fp:
        fp.write('alpha\n\n  beta  \ngamma')
    assert uio.load_text(fp.name) == ['alpha', 'beta', 'gamma']
    udata = importlib.import_module('util.dataset')
    assert udata.load_classes(fp.name) == {'alpha': 1, 'beta': 2, 'gamma': 3}
    os.unlink(fp.name)


def test_frame_statistics_match_reference_counters():
    """ErrorStat / ForegroundF1 (confusion-matrix form) against a direct restatement of util/eval.py:52-85's per-frame rules."""
    sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
    import importlib
    E = importlib.import_module('util.eval')
    rng = np.random.default_rng(0)
    true = rng.integers(0, 5, size=4000) * (rng.random(4000) < 0.3)
    pred = rng.integers(0, 5, size=4000) * (rng.random(4000) < 0.3)
    f1, err = E.ForegroundF1(), E.ErrorStat()
    f1.update(true[:2500], pred[:2500])
    for t, p in zip(true[2500:], pred[2500:]):          # scalar call style of the reference
        f1.update(t, p)
    err.update(true, pred)
    assert err.get() == np.sum(true != pred) / 4000 and err.get_acc() == 1 - err.get()
    tp, fp, fn = {}, {}, {}
    for t, p in zip(true.tolist(), pred.tolist()):
        if p != 0:
            key = 'tp' if t != 0 else 'fp'
            (tp if key == 'tp' else fp)[None] = (tp if key == 'tp' else fp).get(None, 0) + 1
            if p == t:
                tp[p] = tp.get(p, 0) + 1
            else:
                fp[p] = fp.get(p, 0) + 1
                if t != 0:
                    fn[t] = fn.get(t, 0) + 1
        elif t != 0:
            fn[None] = fn.get(None, 0) + 1
            fn[t] = fn.get(t, 0) + 1
    for k in [None, 1, 2, 3, 4, 9]:
        want = (tp.get(k, 0), fp.get(k, 0), fn.get(k, 0))
        assert f1.tp_fp_fn(k) == want
        denom = want[0] + 0.5 * want[1] + 0.5 * want[2]
        assert f1.get(k) == (want[0] / denom if denom else 0.0)


@pytest.mark.needs_reference
def test_unmodified_reference_clis_import_against_the_drop_in():
    """train_tdeed.py:18-26 and evaluate_tdeed_challenge.py:18-25 import, unmodified, with t-deed_b200/ first on sys.path
    (wandb / SoccerNet replaced by the offline stubs), and bind the drop-in's classes / functions."""
    code = (
        'import train_tdeed, evaluate_tdeed_challenge, util.eval, util.score, util.io, model.model, model.modules, model.shift\n'
        'repo = %r\n'
        'for m in (util.eval, util.score, util.io, model.model, model.modules, model.shift):\n'
        '    assert m.__file__.startswith(repo), m.__file__\n'
        'import dataset.frame\n'
        'assert dataset.frame.__file__.startswith("/root/reference")\n'
        'assert train_tdeed.evaluate is util.eval.evaluate and train_tdeed.TDEEDModel is model.model.TDEEDModel\n'
        'assert evaluate_tdeed_challenge.evaluate is util.eval.evaluate\n'
        'for n in ("valMAP_SN", "evaluate_SNB", "aux_evaluate", "label2vector", "predictions2vector"):\n'
        '    f = getattr(util.eval, n)\n'
        '    assert f.__code__.co_filename.startswith("/root/reference"), n\n'
        'assert train_tdeed.valMAP_SN is util.eval.valMAP_SN and train_tdeed.evaluate_SNB is util.eval.evaluate_SNB\n'
        'print("ok")\n') % os.path.join(ROOT, 't-deed_b200')
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 't-deed_b200'), os.path.join(ROOT, 'oracle', 'stubs'),
                                                      '/root/reference']))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, cwd=tempfile.gettempdir())
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stderr[-2000:]


@pytest.mark.needs_reference
def test_reference_names_missing_without_checkout_raise_cleanly():
    code = ('import util.eval\n'
            'try:\n    util.eval.evaluate_SNB\n    print("no error")\n'
            'except ImportError as e:\n    print("ImportError")\n')
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, 't-deed_b200'))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, cwd=tempfile.gettempdir())
    assert r.stdout.strip() == 'ImportError', r.stdout + r.stderr[-1000:]
