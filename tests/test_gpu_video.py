"""Video-level engine (tdeed_b200.pipeline.VideoInference): every unique frame runs stem + s1 + s2 ONCE and the overlapping
clips are assembled from the cached features.  The result must be BIT-IDENTICAL to running each clip of the reference's
clip list (dataset/frame.py:409-423 — starts every clip_len - overlap frames from -pad, zero frames outside the video)
through the whole network and accumulating in clip order (util/eval.py:303-349)."""
from argparse import Namespace

import numpy as np
import pytest
import torch

import tdeed_oracle as O
import postproc_oracle as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda')


def _model(cfg, seed, dev):
    from model.model import TDEEDModel
    args = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=cfg.radi_displacement,
                     feature_arch=cfg.feature_arch, clip_len=cfg.clip_len, n_layers=cfg.n_layers, sgp_ks=cfg.sgp_ks,
                     sgp_r=cfg.sgp_r, num_classes=cfg.num_classes, crop_dim=cfg.crop_dim)
    m = TDEEDModel(device=str(dev), args=args)
    m.load(O.random_state(cfg, seed))
    m._model.eval()
    return m


def _clip(frames, start, clip_len):
    """Zero-padded clip [start, start + clip_len) of a (L,3,H,W) uint8 video (dataset/frame.py:566-625)."""
    out = torch.zeros((clip_len,) + tuple(frames.shape[1:]), dtype=torch.uint8)
    lo, hi = max(start, 0), min(start + clip_len, frames.shape[0])
    if hi > lo:
        out[lo - start:hi - start] = frames[lo:hi]
    return out


def _chunks(videos, n):
    stream = torch.cat(videos)
    for lo in range(0, stream.shape[0], n):
        yield stream[lo:lo + n].contiguous().pin_memory()


@pytest.mark.parametrize('arch,precision,tta,crop,hw', [
    ('rny002_gsf', 'bf16', False, 32, (32, 56)),      # cropped upload path (rows whole, columns cropped)
    ('rny002_gsf', 'bf16', True, None, (32, 48)),     # TTA: plain + flipped view, clip by clip
    ('rny002_gsm', 'fp32', False, 32, (40, 48)),      # exact engine, vertical crop too (no cropped upload)
    ('rny008_gsf', 'bf16', True, None, (64, 64)),
])
def test_video_inference_is_bit_identical_to_per_clip(dev, arch, precision, tta, crop, hw):
    from tdeed_b200.pipeline import VideoInference, VideoScores
    T, overlap, pad = 16, 12, 2
    cfg = O.Config(feature_arch=arch, clip_len=T, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=4, radi_displacement=1, crop_dim=crop)
    m = _model(cfg, 3, dev)
    eng = m._model.engine(precision)
    g = torch.Generator().manual_seed(1)
    lens = [37, 9, 64, 21]                  # incl. a video shorter than a clip and clips that end past the video
    vids = [torch.randint(0, 256, (n, 3) + hw, generator=g, dtype=torch.uint8) for n in lens]
    videos = [('v%d' % i, n, P.clip_starts(n, T, overlap, 1, pad)) for i, n in enumerate(lens)]
    flips = (False, True) if tta else (False,)
    vi = VideoInference(eng, in_hw=hw, clips_per_batch=5, frames_per_chunk=20, flips=flips)
    done = []
    got = vi.run(videos, _chunks(vids, 20), on_video=lambda name, vs: done.append(name))
    assert done == [v[0] for v in videos]
    assert vi.frames_in == sum(lens) and vi.clips_out == sum(len(v[2]) for v in videos)
    K = cfg.num_classes + 1
    for (name, n, starts), frames in zip(videos, vids):
        ref = VideoScores(n, K, dev)
        for s in starts:
            clip = _clip(frames, s, T).unsqueeze(0).to(dev)
            views = []
            for flip in flips:
                _, _, probs = eng.forward(clip, flip=flip)
                views.append(probs.clone())
            ref.add(torch.cat(views), [s] * len(views), tta=tta)
        assert torch.equal(got[name].support, ref.support), name
        assert torch.equal(got[name].scores, ref.scores), '%s: max diff %g' % (name, float((got[name].scores - ref.scores).abs().max()))
    # and a second run over the same engine (graphs replayed, ring reused) reproduces it
    again = vi.run(videos, _chunks(vids, 7))          # pieces that straddle the device buffers
    for name, _, _ in videos:
        assert torch.equal(again[name].scores, got[name].scores)


def test_gather_rows(dev):
    from tdeed_b200 import ops
    g = torch.Generator().manual_seed(0)
    src = torch.randn(13, 7, 5, 8, generator=g).to(torch.bfloat16).to(dev)
    pad = torch.randn(7, 5, 8, generator=g).to(torch.bfloat16).to(dev)
    idx = torch.tensor([3, -1, 12, 0, 0, -1, 7], dtype=torch.int32, device=dev)
    out = ops.gather_rows(src, idx, torch.empty((7, 7, 5, 8), dtype=torch.bfloat16, device=dev), pad_row=pad)
    for i, s in enumerate(idx.tolist()):
        assert torch.equal(out[i], pad if s < 0 else src[s])
    dst = torch.zeros((20, 7, 5, 8), dtype=torch.bfloat16, device=dev)
    didx = torch.tensor([19, 4, 5, 6, 0, 1, 2], dtype=torch.int32, device=dev)
    ops.gather_rows(src, idx, dst, dst_idx=didx, pad_row=pad)
    for s, d in zip(idx.tolist(), didx.tolist()):
        assert torch.equal(dst[d], pad if s < 0 else src[s])
    assert float(dst[10].abs().sum()) == 0.0
    big = torch.randn(3, 100000, generator=g).to(dev)                      # long rows: several slices per row
    o2 = ops.gather_rows(big, torch.tensor([2, 0], dtype=torch.int32, device=dev), torch.empty((2, 100000), device=dev))
    assert torch.equal(o2[0], big[2]) and torch.equal(o2[1], big[0])
