"""`main()`-equivalents of the reference's two scripts on the GPU, with this repo's drop-in modules and nothing else.

/root/reference does not exist on the GPU box, so the scripts themselves cannot run there (tests/test_eval_cpu.py imports them
UNMODIFIED against the drop-in in the build container); these tests replay their call sequences line by line on synthetic data
on disk (a directory of JPEG frames in the fs_comp layout, label JSONs, class.txt):

  train_tdeed.py:89-311             seeds -> datasets / DataLoaders (pin_memory, workers, worker_init_fn) -> TDEEDModel(args)
                                    -> get_optimizer -> ChainedScheduler([LinearLR, CosineAnnealingLR]) -> epochs of
                                    model.epoch(train, optimizer, scaler, lr_scheduler, acc_grad_iter) / model.epoch(val)
                                    -> evaluate(val frames, 'VAL', test=False) (criterion 'map') -> store_json(loss.json)
                                    -> torch.save(state_dict) -> model.load(torch.load(checkpoint_best.pt))
                                    -> evaluate(split data, 'TEST', pred_file, test=True, augment=True)
  evaluate_tdeed_challenge.py:34-101 TDEEDModel(args) -> update_pred_head([13, 18]) -> _num_classes -> model.load(checkpoint)
                                    -> evaluate(challenge data, 'CHALLENGE', pred_file, test=True, augment=False) -> results_spotting.json
"""
import contextlib
import io
import json
import os
import random
from argparse import Namespace

import numpy as np
import pytest
import torch
from torch.optim.lr_scheduler import ChainedScheduler, CosineAnnealingLR, LinearLR
from torch.utils.data import DataLoader, Dataset

import synth_data as S
import tdeed_oracle as O

pytestmark = pytest.mark.gpu


class ClipDataset(Dataset):
    """The batch-dict schema of dataset/frame.py:30-259 (ActionSpotDataset): uint8 'frame' (T,3,H,W), int64 'label' (T),
    optional 'labelD', mixup partner 'frame2' / 'label2' / 'labelD2', 'contains_event'."""

    def __init__(self, video_ds, n, clip_len, radi, mixup, classes):
        self.v, self.n, self.T, self.radi, self.mixup, self.k = video_ds, n, clip_len, radi, mixup, len(classes) + 1

    def __len__(self):
        return self.n

    def _one(self):
        name, length, _ = random.choice(self.v.videos)
        start = random.randint(-2, max(0, length - self.T))
        frames = self.v._frame_reader.load_frames(name, start, start + self.T, pad=True)
        labels = self.v.get_labels(name)
        lab = np.zeros(self.T, np.int64)
        labD = np.zeros(self.T, np.int64)
        for t in range(self.T):
            f = start + t
            for d in range(-self.radi, self.radi + 1):
                if 0 <= f + d < len(labels) and labels[f + d]:
                    lab[t], labD[t] = labels[f + d], d
        return frames, torch.from_numpy(lab), torch.from_numpy(labD)

    def __getitem__(self, unused):
        f, l, d = self._one()
        out = {'frame': f, 'label': l, 'contains_event': int(l.sum() > 0)}
        if self.radi > 0:
            out['labelD'] = d
        if self.mixup:
            f2, l2, d2 = self._one()
            out.update(frame2=f2, label2=l2)
            if self.radi > 0:
                out['labelD2'] = d2
        return out


def test_train_tdeed_main_equivalent(tmp_path):
    from model.model import TDEEDModel
    from util.dataset import load_classes
    from util.eval import evaluate
    from util.io import load_json, store_json
    # ---- data on disk, as the scripts expect it ----
    (tmp_path / 'data').mkdir()
    (tmp_path / 'data' / 'class.txt').write_text('jump\nspin\nstep\nfall\n')
    classes = load_classes(str(tmp_path / 'data' / 'class.txt'))
    assert classes == {'jump': 1, 'spin': 2, 'step': 3, 'fall': 4}
    args = Namespace(model='FineDiving_small', acc_grad_iter=2, seed=1, batch_size=4, clip_len=12, crop_dim=32, dataset='fs_comp',
                     radi_displacement=1, feature_arch='rny002_gsf', learning_rate=8e-4, mixup=True, modality='rgb', num_classes=4,
                     num_epochs=2, warm_up_epochs=1, start_val_epoch=0, temporal_arch='ed_sgp_mixer', n_layers=2, sgp_ks=5, sgp_r=2,
                     criterion='map', num_workers=2, save_dir=str(tmp_path / 'save'), pretrain=None)
    torch.manual_seed(args.seed)                                   # train_tdeed.py:93-95
    np.random.seed(args.seed)
    random.seed(args.seed)
    frames_ds = S.SyntheticVideoDataset(classes, lengths={'va': 40, 'vb': 29}, hw=(32, 56), clip_len=args.clip_len,
                                        overlap_len=args.clip_len // 4 * 3, stride=1, dataset=args.dataset, seed=3, events_per_100=8.0)
    S.write_jpegs(frames_ds, str(tmp_path / 'frames'))            # JPEG decode on the data path, like FrameReaderVideo
    train_data = ClipDataset(frames_ds, 8, args.clip_len, args.radi_displacement, args.mixup, classes)
    val_data = ClipDataset(frames_ds, 4, args.clip_len, args.radi_displacement, False, classes)
    epoch = 0

    def worker_init_fn(id):                                        # train_tdeed.py:126-127
        random.seed(id + epoch * 100)
    loader_batch_size = args.batch_size // args.acc_grad_iter
    train_loader = DataLoader(train_data, shuffle=False, batch_size=loader_batch_size, pin_memory=True, num_workers=args.num_workers,
                              prefetch_factor=2, worker_init_fn=worker_init_fn)
    val_loader = DataLoader(val_data, shuffle=False, batch_size=loader_batch_size, pin_memory=True, num_workers=args.num_workers,
                            prefetch_factor=2, worker_init_fn=worker_init_fn)
    with contextlib.redirect_stdout(io.StringIO()):
        model = TDEEDModel(args=args)                              # device='cuda' default (train_tdeed.py:142)
    optimizer, scaler = model.get_optimizer({'lr': args.learning_rate})
    assert isinstance(optimizer, torch.optim.Optimizer) and scaler is not None
    steps = len(train_loader) // args.acc_grad_iter
    lr_scheduler = ChainedScheduler([LinearLR(optimizer, start_factor=0.01, end_factor=1.0, total_iters=args.warm_up_epochs * steps),
                                     CosineAnnealingLR(optimizer, steps * (args.num_epochs - args.warm_up_epochs))])
    losses, best, before = [], 0, {k: v.clone() for k, v in model.state_dict().items()}
    os.makedirs(args.save_dir, exist_ok=True)
    for epoch in range(args.num_epochs):
        train_loss = model.epoch(train_loader, optimizer, scaler, lr_scheduler=lr_scheduler, acc_grad_iter=args.acc_grad_iter)
        val_loss = model.epoch(val_loader, acc_grad_iter=args.acc_grad_iter)
        with contextlib.redirect_stdout(io.StringIO()):
            val_mAP = evaluate(model, frames_ds, 'VAL', classes, printed=False, test=False)
        assert np.isfinite(train_loss) and np.isfinite(val_loss) and 0.0 <= val_mAP <= 1.0
        losses.append({'epoch': epoch, 'train': train_loss, 'val': val_loss, 'val_mAP': val_mAP})
        store_json(os.path.join(args.save_dir, 'loss.json'), losses, pretty=True)
        if val_mAP >= best:
            best = val_mAP
            torch.save(model.state_dict(), os.path.join(args.save_dir, 'checkpoint_best.pt'))
    assert load_json(os.path.join(args.save_dir, 'loss.json'))[1]['epoch'] == 1
    after = model.state_dict()
    moved = [k for k in before if before[k].dtype.is_floating_point and not torch.equal(before[k], after[k])]
    assert len(moved) > 400                                        # the optimizer stepped every parameter tensor (+ BN statistics)
    assert optimizer.state['flat']['step'] == args.num_epochs * steps   # one fused AdamW step per accumulated batch
    # ---- "START INFERENCE" (train_tdeed.py:236-266) ----
    with contextlib.redirect_stdout(io.StringIO()):
        fresh = TDEEDModel(args=args)
    fresh.load(torch.load(os.path.join(args.save_dir, 'checkpoint_best.pt')))
    split_data = S.SyntheticVideoDataset(classes, lengths={'va': 40, 'vb': 29}, hw=(32, 56), clip_len=args.clip_len,
                                         overlap_len=args.clip_len // 4 * 3, stride=1, dataset=args.dataset, seed=3, events_per_100=8.0,
                                         jpeg_dir=str(tmp_path / 'frames'))
    pred_file = os.path.join(args.save_dir, 'pred-test')
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        mAPs, tolerances = evaluate(fresh, split_data, 'TEST', classes, pred_file, printed=True, test=True,
                                    augment=(args.dataset != 'soccernet') & (args.dataset != 'soccernetball'))
    assert len(mAPs) == len(tolerances) == 3 and 'Results on TEST (w/ SNMS3)' in out.getvalue()
    pred = load_json(pred_file + '.json')
    assert sorted(p['video'] for p in pred) == ['va', 'vb'] and all('events' in p and 'num_events' in p for p in pred)
    # state_dict() / load() round trip: a checkpoint of the model that just trained, loaded into a new model, predicts identically
    torch.save(model.state_dict(), os.path.join(args.save_dir, 'checkpoint_last.pt'))
    with contextlib.redirect_stdout(io.StringIO()):
        clone = TDEEDModel(args=args)
        clone.load(torch.load(os.path.join(args.save_dir, 'checkpoint_last.pt')))
        a = evaluate(model, split_data, 'TEST', classes, os.path.join(args.save_dir, 'pred-a'), printed=True, test=True, augment=True)
        b = evaluate(clone, split_data, 'TEST', classes, os.path.join(args.save_dir, 'pred-b'), printed=True, test=True, augment=True)
    assert [float(x) for x in a[0]] == [float(x) for x in b[0]]
    assert load_json(os.path.join(args.save_dir, 'pred-a.json')) == load_json(os.path.join(args.save_dir, 'pred-b.json'))


def test_evaluate_tdeed_challenge_main_equivalent(tmp_path):
    from model.model import TDEEDModel
    from util.eval import evaluate
    from util.io import load_json
    classes = {'c%02d' % i: i for i in range(1, 13)}               # 12 SoccerNetBall classes
    pretrain_classes = {'p%02d' % i: i for i in range(1, 18)}      # 17 SoccerNet classes (args.pretrain)
    args = Namespace(model='SoccerNetBall_challenge1', seed=1, batch_size=4, acc_grad_iter=1, clip_len=12, crop_dim=None,
                     dataset='soccernetball', radi_displacement=4, feature_arch='rny002_gsf', modality='rgb', num_classes=12,
                     temporal_arch='ed_sgp_mixer', n_layers=2, sgp_ks=9, sgp_r=4, save_dir=str(tmp_path / 'save'),
                     pretrain={'dataset': 'soccernet', 'num_classes': 17})
    cfg = O.Config(feature_arch=args.feature_arch, clip_len=args.clip_len, n_layers=2, sgp_ks=9, sgp_r=4, num_classes=12,
                   radi_displacement=4, crop_dim=None, double_head=[13, 18])
    os.makedirs(tmp_path / 'checkpoints', exist_ok=True)
    torch.save(O.random_state(cfg, 17), tmp_path / 'checkpoints' / 'checkpoint_best.pt')   # "a checkpoint saved from random_state"
    with contextlib.redirect_stdout(io.StringIO()):
        model = TDEEDModel(args=args)                               # evaluate_tdeed_challenge.py:65
    n_classes = [len(classes) + 1, len(pretrain_classes) + 1]
    model._model.update_pred_head(n_classes)                        # :68-73
    model._num_classes = np.array(n_classes).sum()
    model.load(torch.load(tmp_path / 'checkpoints' / 'checkpoint_best.pt'))
    split_data = S.SyntheticVideoDataset(classes, lengths={'league/game_a': 61, 'league/game_b': 44}, hw=(32, 48),
                                         clip_len=args.clip_len, overlap_len=args.clip_len // 4 * 3, stride=2, dataset=args.dataset, seed=9)
    pred_file = os.path.join(args.save_dir, 'pred-challenge')
    os.makedirs(args.save_dir, exist_ok=True)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        mAPs, tolerances = evaluate(model, split_data, 'CHALLENGE', classes, pred_file, printed=True, test=True,
                                    augment=(args.dataset != 'soccernet') & (args.dataset != 'soccernetball'))
    assert mAPs is None and tolerances is None and 'Storing predictions Challenge with SNMS' in out.getvalue()
    for game in ('league/game_a', 'league/game_b'):
        res = load_json(os.path.join(args.save_dir, 'preds', game, 'results_spotting.json'))
        assert res['UrlLocal'] == game and isinstance(res['predictions'], list)
        for p in res['predictions'][:50]:
            assert set(p) == {'gameTime', 'label', 'position', 'confidence', 'half'} and p['half'] == 1 and p['label'] in classes
            assert abs(p['position'] - 80 * round(p['position'] / 80)) <= 1 and 0.01 <= p['confidence'] <= 1.0   # int(frame / 25 * 1000 * stride)
    total = sum(len(load_json(os.path.join(args.save_dir, 'preds', g, 'results_spotting.json'))['predictions'])
                for g in ('league/game_a', 'league/game_b'))
    assert total > 0
