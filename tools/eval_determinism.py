#!/usr/bin/env python3
"""Determinism check of the drop-in evaluation path on the bench's synthetic dataset: the five videos hold the same frames, so their
score tables must be equal, and two runs must agree bit for bit.  Compares the threaded frame source with plain in-order pieces."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
import torch
import bench
from model.model import TDEEDModel
import util.eval as E

with contextlib.redirect_stdout(io.StringIO()):
    model = TDEEDModel(device='cuda:0', args=bench.model_args())
bench.randomize_(model._model, 0)
g = torch.Generator().manual_seed(7)
frames = torch.randint(0, 256, (bench.VIDEO_FRAMES, 3, bench.FRAME_H, bench.FRAME_W), generator=g, dtype=torch.uint8)
ds = bench.SyntheticVideoDataset(frames, ['video%02d' % i for i in range(5)])
dev = torch.device('cuda:0')


def run(augment):
    with contextlib.redirect_stderr(io.StringIO()):
        sc = E._stream_scores(model, ds, None, augment, 5, dev)
    torch.cuda.synchronize()
    return {k: (v.scores.clone(), v.support.clone()) for k, v in sc.items()}


def diff(a, b, what):
    bad = 0
    for k in a:
        for i, nm in ((0, 'scores'), (1, 'support')):
            if not torch.equal(a[k][i], b[k][i]):
                d = (a[k][i].float() - b[k][i].float()).abs()
                rows = torch.nonzero(d.reshape(d.shape[0], -1).amax(1) > 0).flatten()
                print('%s: %s %s differs in %d frames [%d .. %d], max |d| %.3g' % (what, k, nm, rows.numel(), int(rows[0]), int(rows[-1]), float(d.max())))
                bad += 1
    return bad


for augment in (False, True):
    r1 = run(augment)
    r2 = run(augment)
    r3 = run(augment)
    bad = diff(r1, r2, 'run1 vs run2 (augment=%s)' % augment) + diff(r2, r3, 'run2 vs run3 (augment=%s)' % augment)
    v0 = r1['video00']
    for k in r1:
        bad += diff({k: r1[k]}, {k: v0}, 'within run1 (augment=%s), vs video00' % augment)
    print('augment=%s: %s' % (augment, 'DETERMINISTIC, all videos equal' if bad == 0 else '%d mismatches' % bad))
