"""Bit-exact parity of the post-processing kernels (clip accumulation, event extraction, NMS, soft-NMS)
against the golden vectors from the unmodified reference and against the CPU oracle on random inputs."""
import os

import numpy as np
import pytest
import torch

import postproc_oracle as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def _cases(g):
    return sorted({k.split('_')[0] for k in g.files})


def run_video(dev, preds, starts, length, k, batch, thr=0.01, mode=0):
    from tdeed_b200 import ops
    scores = torch.zeros(length, k, device=dev)
    support = torch.zeros(length, dtype=torch.int32, device=dev)
    for i in range(0, len(starts), batch):
        ops.clip_accumulate(scores, support, torch.as_tensor(preds[i:i + batch]).to(dev).contiguous(),
                            torch.as_tensor(np.asarray(starts[i:i + batch], np.int32)).to(dev), mode)
    raw = scores.clone(), support.clone()
    ev = ops.extract_events(scores, support, thr)
    return raw, scores, support, ev


def nms_gpu(ev, k, window, thr, soft):
    from tdeed_b200 import ops
    of, ol, os_, oc = ops.nms(ev['hr_frame'], ev['hr_label'], ev['hr_score'], ev['counts'][1:2], k, window, thr, soft)
    n = int(oc.item())
    return of[:n].cpu().numpy(), ol[:n].cpu().numpy(), os_[:n].cpu().numpy()


@pytest.mark.parametrize('batch', [4, 1, 7])
def test_golden_postproc(golden_dir, dev, batch):
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    for c in _cases(g):
        length, k, w0, w1 = [int(v) for v in g[c + '_meta']]
        thr = float(g[c + '_nms_thr'])
        (raw_s, raw_sup), scores, support, ev = run_video(dev, g[c + '_preds'], g[c + '_starts'].tolist(), length, k, batch)
        assert np.array_equal(raw_s.cpu().numpy(), g[c + '_scores_sum'])
        assert np.array_equal(raw_sup.cpu().numpy(), g[c + '_support'])
        assert np.array_equal(scores.cpu().numpy(), g[c + '_scores_norm'])
        n_ev, n_hr = ev['counts'].cpu().tolist()
        for tag, n in (('ev', n_ev), ('hr', n_hr)):
            assert np.array_equal(ev[tag + '_frame'][:n].cpu().numpy(), g[c + '_%s_frame' % tag])
            assert np.array_equal(ev[tag + '_label'][:n].cpu().numpy(), g[c + '_%s_label' % tag])
            assert np.array_equal(ev[tag + '_score'][:n].cpu().numpy().astype(np.float64), g[c + '_%s_score' % tag])
        f, l, s = nms_gpu(ev, k, w0, thr, soft=False)
        assert np.array_equal(f, g[c + '_nms_frame']) and np.array_equal(l, g[c + '_nms_label'])
        assert np.array_equal(s, g[c + '_nms_score'])
        f, l, s = nms_gpu(ev, k, w1, 0.01, soft=True)
        assert np.array_equal(f, g[c + '_snms_frame']) and np.array_equal(l, g[c + '_snms_label'])
        ref_s = g[c + '_snms_score']
        bad = np.nonzero(s != ref_s)[0]
        assert len(bad) == 0, [(int(f[i]), int(l[i]), float(s[i]).hex(), float(ref_s[i]).hex()) for i in bad[:8]]   # float64 bit-exact


def test_nms_random_vs_oracle(dev):
    from tdeed_b200 import ops
    rng = np.random.default_rng(11)
    for trial in range(30):
        length, k = int(rng.integers(1, 400)), int(rng.integers(2, 20))
        vals = np.asarray([0.0, 0.0, 0.0, 0.005, 0.01, 0.0100001, 0.2, 0.2, 0.5, 0.7], np.float32)
        scores = rng.choice(vals, size=(length, k)).astype(np.float32)      # many exact ties and threshold edges
        scores[:, 0] = 0.3
        support = rng.integers(0, 4, length).astype(np.int32)
        w = int(rng.integers(0, 9))
        thr = float(rng.choice([0.0, 0.01, 0.1]))
        s_ref, sup_ref = scores.copy(), support.copy()
        pred_ref, ev_ref, hr_ref = P.frame_predictions(s_ref, sup_ref, 0.01)
        ds, dsup = torch.as_tensor(scores).to(dev), torch.as_tensor(support).to(dev)
        ev = ops.extract_events(ds, dsup, 0.01)
        assert np.array_equal(ds.cpu().numpy(), s_ref) and np.array_equal(ev['pred'].cpu().numpy(), pred_ref)
        n_ev, n_hr = ev['counts'].cpu().tolist()
        assert n_hr == len(hr_ref[0]) and n_ev == len(ev_ref[0])
        assert np.array_equal(ev['hr_frame'][:n_hr].cpu().numpy(), hr_ref[0])
        assert np.array_equal(ev['hr_label'][:n_hr].cpu().numpy(), hr_ref[1])
        if n_hr == 0:
            continue
        for soft in (False, True):
            ref = P.soft_nms(*hr_ref, window=w, threshold=0.01) if soft else P.nms(*hr_ref, window=w, threshold=thr)
            f, l, s = nms_gpu(ev, k, w, 0.01 if soft else thr, soft)
            assert np.array_equal(f, ref[0]) and np.array_equal(l, ref[1]), (trial, soft)
            assert np.array_equal(s, ref[2].astype(np.float64)), (trial, soft)


def test_tta_accumulate_and_full_size_properties(dev):
    """SoccerNetBall-sized video (71 594 frames, 13 classes): support equals clip coverage, scores are a convex
    combination (max <= 1), NMS output is sorted, respects the window per label and is idempotent."""
    from tdeed_b200 import ops
    length, k, T = 71594, 13, 100
    starts = P.clip_starts(143188, 100, 75, 2)
    g = torch.Generator(device='cpu').manual_seed(0)
    scores = torch.zeros(length, k, device=dev)
    support = torch.zeros(length, dtype=torch.int32, device=dev)
    for i in range(0, len(starts), 64):
        st = starts[i:i + 64]
        pred = torch.softmax(torch.randn(len(st), T, k, generator=g) * 3, dim=2).to(dev)
        ops.clip_accumulate(scores, support, pred, torch.as_tensor(np.asarray(st, np.int32)).to(dev), 1)
    cover = np.zeros(length, np.int64)
    for s in starts:
        cover[max(s, 0):min(s + T, length)] += 1
    assert np.array_equal(support.cpu().numpy(), cover)
    ev = ops.extract_events(scores, support, 0.01)
    assert float(scores.max()) <= 1.0 + 1e-6 and abs(float(scores.sum(1).mean()) - 1.0) < 1e-4
    f, l, s = nms_gpu(ev, k, 6, 0.01, soft=False)
    assert np.all(np.diff(f) >= 0)
    for lab in range(1, k):
        fl = f[l == lab]
        assert np.all(np.diff(fl) > 6)
    # idempotence: NMS of an NMS output keeps every event (same-frame tie order follows the NEW list's label
    # first-appearance order, exactly like the reference, so compare as sets)
    of = torch.as_tensor(f).to(dev); ol = torch.as_tensor(l).to(dev); osc = torch.as_tensor(s.astype(np.float32)).to(dev)
    cnt = torch.tensor([len(f)], dtype=torch.int32, device=dev)
    f2, l2, s2, c2 = ops.nms(of, ol, osc, cnt, k, 6, 0.01, False)
    n2 = int(c2.item())
    assert n2 == len(f) and np.array_equal(f2[:n2].cpu().numpy(), f)
    assert set(zip(f2[:n2].cpu().tolist(), l2[:n2].cpu().tolist())) == set(zip(f.tolist(), l.tolist()))
    # and the full-size result equals the CPU oracle on the same event list (bit-exact)
    nh = int(ev['counts'][1].item())
    hr = (ev['hr_frame'][:nh].cpu().numpy(), ev['hr_label'][:nh].cpu().numpy(), ev['hr_score'][:nh].cpu().numpy())
    rf, rl, rs = P.nms(*hr, window=6, threshold=0.01)
    assert np.array_equal(f, rf) and np.array_equal(l, rl) and np.array_equal(s, rs.astype(np.float64))


class _FakeDataset:
    def __init__(self, videos):
        self.videos = videos


def test_util_eval_dropin_matches_reference_golden(golden_dir, dev):
    """The drop-in util.eval functions (reference signatures, list-of-dict wire format) against the golden vectors
    produced by the reference's own process_frame_predictions_challenge / (soft_)non_maximum_supression."""
    import importlib
    ue = importlib.import_module('util.eval')
    assert 't-deed_b200' in ue.__file__
    g = np.load(os.path.join(golden_dir, 'postproc.npz'))
    for c in _cases(g):
        length, k, w0, w1 = [int(v) for v in g[c + '_meta']]
        thr = float(g[c + '_nms_thr'])
        classes = {'c%d' % j: j for j in range(1, k)}
        scores, support = g[c + '_scores_sum'].copy(), g[c + '_support'].copy()
        ds = _FakeDataset([('vid', length, 25.0)])
        pe, pehr, ps = ue.process_frame_predictions_challenge(ds, classes, {'vid': (scores, support)}, 0.01)
        assert np.array_equal(scores, g[c + '_scores_norm'])                 # normalised in place like the reference
        for tag, vp in (('ev', pe), ('hr', pehr)):
            f, l, s = P.from_dicts(vp[0], classes)
            assert np.array_equal(f, g[c + '_%s_frame' % tag]) and np.array_equal(l, g[c + '_%s_label' % tag])
            assert np.array_equal(s, g[c + '_%s_score' % tag])
        for tag, out in (('nms', ue.non_maximum_supression(pehr, window=w0, threshold=thr)),
                         ('snms', ue.soft_non_maximum_supression(pehr, window=w1, threshold=0.01))):
            f, l, s = P.from_dicts(out[0], classes)
            assert out[0]['num_events'] == len(f) and out[0]['video'] == 'vid'
            assert np.array_equal(f, g[c + '_%s_frame' % tag]) and np.array_equal(l, g[c + '_%s_label' % tag])
            assert np.array_equal(s, g[c + '_%s_score' % tag])
