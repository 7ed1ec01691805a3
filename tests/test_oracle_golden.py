"""Pins the CPU oracle (oracle/tdeed_oracle.py, oracle/postproc_oracle.py) against the golden vectors
that oracle/gen_golden.py produced by running the UNMODIFIED reference (tests/golden/*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch

import tdeed_oracle as O
import postproc_oracle as P
from gen_golden import MODEL_CASES, weights_digest


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize('name', sorted(MODEL_CASES))
def test_model_oracle_matches_reference_golden(name, golden_dir):
    kw, _, wseed, _ = MODEL_CASES[name]
    g = np.load(os.path.join(golden_dir, 'model_%s.npz' % name))
    cfg = O.Config(**kw)
    sd = O.random_state(cfg, wseed)
    np.testing.assert_allclose(weights_digest(sd), g['weights_digest'], rtol=1e-12)   # RNG did not drift
    frames = torch.from_numpy(g['frames'])
    for flip in (False, True):
        sfx = '_flip' if flip else ''
        with torch.no_grad():
            logits, displ = O.forward(sd, cfg, frames, flip=flip)
        assert rel_err(logits.numpy(), g['logits' + sfx]) < 2e-5
        if displ is not None:
            assert rel_err(displ.numpy(), g['displ' + sfx]) < 2e-5
        cls, probs = O.predict(sd, cfg, frames, flip=flip)
        assert np.abs(probs - g['probs' + sfx]).max() < 1e-5
        # argmax may only differ where the two top probabilities are within the fp noise
        diff = cls != g['cls' + sfx]
        if diff.any():
            top2 = np.sort(g['probs' + sfx], axis=2)[..., -2:]
            assert np.all((top2[..., 1] - top2[..., 0])[diff] < 1e-5)


def _cases(g):
    return sorted({k.split('_')[0] for k in g.files})


def test_postproc_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    for c in _cases(g):
        length, k, w0, w1 = [int(v) for v in g[c + '_meta']]
        thr = float(g[c + '_nms_thr'])
        scores = np.zeros((length, k), np.float32)
        support = np.zeros(length, np.int32)
        for s, p in zip(g[c + '_starts'], g[c + '_preds']):
            P.accumulate_batched(scores, support, p.copy(), int(s))
        assert np.array_equal(scores, g[c + '_scores_sum'])
        assert np.array_equal(support, g[c + '_support'])
        pred, ev, hr = P.frame_predictions(scores, support, 0.01)
        assert np.array_equal(scores, g[c + '_scores_norm'])
        for tag, (f, l, s) in (('ev', ev), ('hr', hr)):
            assert np.array_equal(f, g[c + '_%s_frame' % tag])
            assert np.array_equal(l, g[c + '_%s_label' % tag])
            assert np.array_equal(s.astype(np.float64), g[c + '_%s_score' % tag])
        f, l, s = P.nms(*hr, window=w0, threshold=thr)
        assert np.array_equal(f, g[c + '_nms_frame']) and np.array_equal(l, g[c + '_nms_label'])
        assert np.array_equal(s.astype(np.float64), g[c + '_nms_score'])
        f, l, s = P.soft_nms(*hr, window=w1, threshold=0.01)
        assert np.array_equal(f, g[c + '_snms_frame']) and np.array_equal(l, g[c + '_snms_label'])
        assert np.array_equal(s, g[c + '_snms_score'])          # float64, bit-exact
