#!/usr/bin/env python3
"""Phase timing of util.eval.evaluate on the bench's synthetic dataset (TDEED_EVAL_TIMING=1)."""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 't-deed_b200'))
os.environ['TDEED_EVAL_TIMING'] = '1'
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
classes = {'c%d' % i: i for i in range(1, 5)}
for rep in range(3):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        E.evaluate(model, ds, 'VAL', classes, printed=False, test=False, augment=False)
    print('evaluate total %.1f ms' % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
# source alone
from tdeed_b200.pipeline import ThreadedFrameSource
pieces = E._FramePieces(ds, [(v, n, (v, -5)) for v, n, _ in ds.videos], E.STREAM_PIECE_FRAMES)
st = torch.cuda.Stream()
t0 = time.perf_counter()
n = 0
for p in ThreadedFrameSource(pieces, st, workers=8):
    n += p.shape[0]
print('frame source alone: %d frames in %.1f ms' % (n, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
