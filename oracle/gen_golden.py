#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) on seeded synthetic inputs.

Run in the build container only:   python oracle/gen_golden.py
The fixtures are committed; the GPU box (no /root/reference) only reads them.

Weights come from oracle.tdeed_oracle.random_state(cfg, seed) and are loaded into the reference
model with load_state_dict(strict=True) — which also pins the oracle's state-dict layout
(names, order, shapes) against the reference's.
"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import tdeed_oracle as O            # noqa: E402
import postproc_oracle as P         # noqa: E402
from ref_import import reference_modules  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

# name -> (oracle Config kwargs, input (B,T,H,W), weight seed, input seed)
MODEL_CASES = {
    'rny002_gsf_displ': (dict(feature_arch='rny002_gsf', clip_len=16, n_layers=2, sgp_ks=7, sgp_r=4,
                              num_classes=4, radi_displacement=2, crop_dim=64), (2, 16, 64, 80), 1, 11),
    'rny002_gsf_l3_ks5': (dict(feature_arch='rny002_gsf', clip_len=20, n_layers=3, sgp_ks=5, sgp_r=4,
                               num_classes=4, radi_displacement=1, crop_dim=64), (1, 20, 64, 114), 2, 12),
    'rny002_gsm_nodispl': (dict(feature_arch='rny002_gsm', clip_len=12, n_layers=2, sgp_ks=9, sgp_r=2,
                                num_classes=6, radi_displacement=0, crop_dim=None), (1, 12, 64, 96), 3, 13),
    'rny008_gsf_double': (dict(feature_arch='rny008_gsf', clip_len=10, n_layers=2, sgp_ks=9, sgp_r=4,
                               num_classes=12, radi_displacement=4, crop_dim=None, double_head=[13, 18]),
                          (1, 10, 64, 96), 4, 14),
}


def ref_args(cfg):
    return Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=cfg.radi_displacement,
                     feature_arch=cfg.feature_arch, clip_len=cfg.clip_len, n_layers=cfg.n_layers,
                     sgp_ks=cfg.sgp_ks, sgp_r=cfg.sgp_r, num_classes=cfg.num_classes, crop_dim=cfg.crop_dim)


def build_reference_model(mods, cfg, sd):
    """Reference TDEEDModel on CPU with the oracle-generated weights (strict load)."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = mods['model.model'].TDEEDModel(device='cpu', args=ref_args(cfg))
    if cfg.double_head:
        # update_pred_head hard-codes .cuda() (model/model.py:169-172) -> do the same by hand on CPU
        model._model._pred_fine = mods['model.modules'].FC2Layers(cfg.feat_dim, cfg.double_head)
        model._model._double_head = True
        model._num_classes = sum(cfg.double_head)
    if cfg.shift_mode == 'gsm':
        # _GSM builds its zero pad with torch.cuda.FloatTensor (gsm.py:67,84,87): CPU stand-in
        mods['model.impl.gsm'].ftens = lambda *s: torch.zeros(*s)
    model.load(sd)      # strict
    return model


def weights_digest(sd):
    return np.asarray([float(v.double().abs().sum()) for k, v in sd.items() if v.dtype.is_floating_point][:64])


def gen_models(mods):
    for name, (kw, (b, t, h, w), wseed, iseed) in MODEL_CASES.items():
        cfg = O.Config(**kw)
        sd = O.random_state(cfg, wseed)
        model = build_reference_model(mods, cfg, sd)
        g = torch.Generator().manual_seed(iseed)
        frames = torch.randint(0, 256, (b, t, 3, h, w), generator=g, dtype=torch.uint8)
        out = {'frames': frames.numpy(), 'weight_seed': wseed, 'weights_digest': weights_digest(sd)}
        model._model.eval()
        with torch.no_grad():
            for flip in (False, True):
                pred, _ = model._model(frames.float(), inference=True, augment_inference=flip)
                if isinstance(pred, dict):
                    logits, displ = pred['im_feat'], pred['displ_feat']
                else:
                    logits, displ = pred, None
                sfx = '_flip' if flip else ''
                out['logits' + sfx] = logits.numpy()
                if displ is not None:
                    out['displ' + sfx] = displ.numpy()
                cls, probs = model.predict(frames, use_amp=False, augment_inference=flip)
                out['cls' + sfx] = cls
                out['probs' + sfx] = probs
        np.savez_compressed(os.path.join(GOLDEN, 'model_%s.npz' % name), **out)
        print(name, 'logits', out['logits'].shape, 'absmax', float(np.abs(out['logits']).max()),
              'probs max', float(out['probs'].max()))


class _FakeDataset:
    def __init__(self, videos):
        self.videos = videos


def synth_scores(rng, length, k, smooth=5, temp=2.0):
    """Smoothed softmax(N(0, temp^2)) so that a few % of (frame, class) cells exceed 0.01."""
    z = rng.normal(0, temp, size=(length + smooth - 1, k)).astype(np.float32)
    z[:, 0] += 3.0
    kern = np.ones(smooth, np.float32) / smooth
    z = np.stack([np.convolve(z[:, j], kern, mode='valid') for j in range(k)], axis=1) * np.float32(2.5)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def gen_postproc(mods):
    ev = mods['util.eval']
    rng = np.random.default_rng(7)
    cases = {}
    for ci, (length, k, windows, nms_thr) in enumerate([(300, 5, (1, 3), 0.01), (1000, 13, (6, 12), 0.01),
                                                         (257, 7, (3, 6), 0.10), (40, 5, (1, 3), 0.01)]):
        classes = {'c%d' % j: j for j in range(1, k)}
        clip_len, overlap = 100, 75
        starts = P.clip_starts(length, clip_len, overlap, 1)
        preds = [synth_scores(rng, clip_len, k) for _ in starts]
        # some all-zero rows (frames never hit by the displacement scatter, modules.py:406-414)
        for p in preds:
            p[rng.random(clip_len) < 0.1] = 0
        scores = np.zeros((length, k), np.float32)
        support = np.zeros(length, np.int32)
        for s, p in zip(starts, preds):
            P.accumulate_batched(scores, support, p, s)      # oracle accumulate ...
        # ... cross-checked against an inline transcription-free use of the reference loop is not
        # possible (it lives inside evaluate()); the numeric part is `+=` in clip order.
        pred_dict = {'vid': (scores.copy(), support.copy())}
        ds = _FakeDataset([('vid', length, 25.0)])
        pe, pehr, _ = ev.process_frame_predictions_challenge(ds, classes, pred_dict, high_recall_score_threshold=0.01)
        nms = ev.non_maximum_supression(pehr, window=windows[0], threshold=nms_thr)
        snms = ev.soft_non_maximum_supression(pehr, window=windows[1], threshold=0.01)

        def pack(vp):
            f, l, s = P.from_dicts(vp[0], classes)
            return f, l, s
        c = 'case%d_' % ci
        cases[c + 'meta'] = np.asarray([length, k, windows[0], windows[1]], np.int64)
        cases[c + 'nms_thr'] = np.asarray(nms_thr)
        cases[c + 'starts'] = np.asarray(starts, np.int64)
        cases[c + 'preds'] = np.stack(preds)
        cases[c + 'scores_sum'] = scores
        cases[c + 'support'] = support
        cases[c + 'scores_norm'] = pred_dict['vid'][0]
        for tag, vp in (('ev', pe), ('hr', pehr), ('nms', nms), ('snms', snms)):
            f, l, s = pack(vp)
            cases[c + tag + '_frame'] = f
            cases[c + tag + '_label'] = l
            cases[c + tag + '_score'] = s
        print('postproc case', ci, 'hr events', len(pehr[0]['events']), 'nms', len(nms[0]['events']),
              'snms', len(snms[0]['events']))
    np.savez_compressed(os.path.join(GOLDEN, 'postproc.npz'), **cases)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    torch.backends.mkldnn.enabled = True
    with reference_modules() as mods:
        gen_models(mods)
        gen_postproc(mods)


if __name__ == '__main__':
    main()
