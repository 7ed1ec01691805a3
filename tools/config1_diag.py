#!/usr/bin/env python3
"""Why can the post-processed probabilities of the bf16 engine differ a lot from fp32 at single frames?  Raw logits / displacement of
the FineDiving_small forward (B = 4) in both precisions, and the frames whose rounded displacement flips."""
import contextlib, io, os, sys
from argparse import Namespace
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 't-deed_b200'), os.path.join(ROOT, 'oracle'), ROOT):
    sys.path.insert(0, p)
import tdeed_oracle as O
from model.model import TDEEDModel
cfg = O.named_config('FineDiving_small', feature_arch='rny002_gsf')
sd = O.random_state(cfg, 0)
margs = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=2, feature_arch='rny002_gsf', clip_len=100,
                  n_layers=2, sgp_ks=7, sgp_r=4, num_classes=4, crop_dim=224)
with contextlib.redirect_stdout(io.StringIO()):
    m = TDEEDModel(device='cuda', args=margs)
m.load(sd)
torch.manual_seed(0)
x = torch.randint(0, 256, (4, 100, 3, 224, 224), dtype=torch.uint8).cuda()
out = {}
for prec in ('fp32', 'bf16'):
    eng = m._model.engine(prec)
    logits, displ, probs = eng.forward(x)
    out[prec] = (logits.float().cpu().numpy().copy(), displ.float().cpu().numpy().copy(), probs.float().cpu().numpy().copy())
lf, df, pf = out['fp32']
lb, db, pb = out['bf16']
print('logits: max |bf16 - fp32| = %.4g (max |fp32| = %.4g)' % (np.abs(lb - lf).max(), np.abs(lf).max()))
print('displ : max |bf16 - fp32| = %.4g (max |fp32| = %.4g)' % (np.abs(db - df).max(), np.abs(df).max()))
flip = np.argwhere(np.rint(db) != np.rint(df))
print('frames whose rounded displacement differs: %d of %d' % (len(flip), df.size))
for idx in flip[:8]:
    i = tuple(idx)
    print('  clip %d frame %d: displ fp32 %.5f bf16 %.5f' % (i[0], i[1], df[i], db[i]))
same = np.rint(db) == np.rint(df)
print('probs : max |bf16 - fp32| over all frames = %.4g' % np.abs(pb - pf).max())
# one-clip-at-a-time equals the batch (bit for bit)
eng = m._model.engine('bf16')
one = torch.cat([eng.forward(x[i:i + 1])[0].float().cpu() for i in range(4)])
print('batch == clips one by one (bf16 logits):', bool(torch.equal(one, torch.from_numpy(lb))))
