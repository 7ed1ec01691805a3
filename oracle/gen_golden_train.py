#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — generate tests/golden/train_*.npz by running the UNMODIFIED reference training path
(`TDEEDModel.epoch` with an optimizer, /root/reference/model/model.py:193-332) on CPU for ONE batch.

The optimizer handed to `epoch` only records: its step()/zero_grad() are no-ops, so after the call every
parameter still carries the gradient the reference computed.  As SURVEY.md §8c prescribes for gradient parity the
random augmentation pipeline is replaced by Identity and Dropout p is set to 0 (both are sampled per call and have
no deterministic counterpart); the crop is the identity (crop_dim == frame size or None).

Run in the build container only:   python oracle/gen_golden_train.py
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import tdeed_oracle as O            # noqa: E402
from ref_import import reference_modules  # noqa: E402
from gen_golden import build_reference_model, GOLDEN  # noqa: E402

# name -> (Config kwargs, (B,T,H,W), weight seed, input seed, mixup)
TRAIN_CASES = {
    'rny002_gsf_displ': (dict(feature_arch='rny002_gsf', clip_len=12, n_layers=2, sgp_ks=5, sgp_r=4, num_classes=4,
                              radi_displacement=2, crop_dim=64), (2, 12, 64, 64), 21, 31, False),
    'rny002_gsf_mixup': (dict(feature_arch='rny002_gsf', clip_len=10, n_layers=2, sgp_ks=7, sgp_r=2, num_classes=5,
                              radi_displacement=1, crop_dim=None), (2, 10, 64, 96), 22, 32, True),
    'rny008_gsf_nodispl': (dict(feature_arch='rny008_gsf', clip_len=8, n_layers=3, sgp_ks=9, sgp_r=4, num_classes=7,
                                radi_displacement=0, crop_dim=None), (2, 8, 64, 64), 23, 33, True),
}
# gradients of these tensors are stored in full; of all the others only (sum, abs-sum, L2 norm)
FULL_GRADS = ('temp_enc', '_features.stem.conv.weight', '_features.stem.bn.weight', '_features.s1.b1.conv2.conv.weight',
              '_features.s1.b1.se.fc1.weight', '_features.s3.b1.conv1.gs.conv3D.weight', '_features.s3.b1.conv1.gs.bn.weight',
              '_features.s3.b1.conv1.gs.channel_conv1.weight', '_features.s3.b2.conv1.net.conv.weight',
              '_features.s4.b1.downsample.conv.weight', '_features.s4.b1.conv3.bn.bias',
              '_temp_fine._sgp.0.ln.weight', '_temp_fine._sgp.0.convkw.weight', '_temp_fine._sgp.1.global_fc.weight',
              '_temp_fine._sgp.2.gn.weight', '_temp_fine._sgpMixer.0.psi2.weight', '_temp_fine._sgpMixer.0.ln2.bias',
              '_temp_fine._sgpMixer.1.global_fc1.bias', '_pred_fine._fc_out.weight', '_pred_displ._fc_out.weight')


class RecordingOptimizer:
    def zero_grad(self):
        pass

    def step(self):
        pass


def make_batch(cfg, shape, seed, mixup):
    b, t, h, w = shape
    g = torch.Generator().manual_seed(seed)
    k = cfg.num_classes + 1

    def labels():
        lab = torch.zeros((b, t), dtype=torch.int64)
        hit = torch.rand((b, t), generator=g) < 0.3
        lab[hit] = torch.randint(1, k, (int(hit.sum()),), generator=g)
        return lab

    batch = {'frame': torch.randint(0, 256, (b, t, 3, h, w), generator=g, dtype=torch.uint8), 'label': labels()}
    if cfg.radi_displacement > 0:
        batch['labelD'] = torch.randint(-cfg.radi_displacement, cfg.radi_displacement + 1, (b, t), generator=g)
    if mixup:
        batch['frame2'] = torch.randint(0, 256, (b, t, 3, h, w), generator=g, dtype=torch.uint8)
        batch['label2'] = labels()
        if cfg.radi_displacement > 0:
            batch['labelD2'] = torch.randint(-cfg.radi_displacement, cfg.radi_displacement + 1, (b, t), generator=g)
    return batch


def main():
    with reference_modules() as mods:
        for name, (kw, shape, wseed, iseed, mix) in TRAIN_CASES.items():
            cfg = O.Config(**kw)
            sd = O.random_state(cfg, wseed)
            model = build_reference_model(mods, cfg, sd)
            net = model._model
            net.augmentation = torch.nn.Identity()
            for m in net.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            batch = make_batch(cfg, shape, iseed, mix)
            random.seed(iseed)
            lam = [random.betavariate(0.2, 0.2) for _ in range(shape[0])] if mix else []
            random.seed(iseed)                         # epoch() draws the same lambdas
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                loss = model.epoch([batch], optimizer=RecordingOptimizer(), scaler=None, fg_weight=5)
            out = {'loss': np.float64(loss), 'lam': np.asarray(lam, np.float64), 'weight_seed': wseed}
            for k_, v in batch.items():
                out['batch_' + k_] = v.numpy()
            names, stats = [], []
            for pname, p in net.named_parameters():
                gnp = p.grad.detach().double().numpy()
                names.append(pname)
                stats.append([gnp.sum(), np.abs(gnp).sum(), np.sqrt((gnp * gnp).sum())])
                if pname in FULL_GRADS:
                    out['grad/' + pname] = p.grad.detach().numpy()
            out['grad_names'] = np.asarray(names)
            out['grad_stats'] = np.asarray(stats, np.float64)
            after = net.state_dict()
            for bn in ('_features.stem.bn', '_features.s2.b1.conv3.bn', '_features.s3.b1.conv1.gs.bn', '_features.s4.b1.downsample.bn'):
                out['after/' + bn + '.running_mean'] = after[bn + '.running_mean'].numpy()
                out['after/' + bn + '.running_var'] = after[bn + '.running_var'].numpy()
                out['after/' + bn + '.num_batches_tracked'] = after[bn + '.num_batches_tracked'].numpy()
            np.savez_compressed(os.path.join(GOLDEN, 'train_%s.npz' % name), **out)
            print(name, 'loss', loss, 'params', len(names), 'max |grad| L2', float(np.max(out['grad_stats'][:, 2])))


if __name__ == '__main__':
    main()
