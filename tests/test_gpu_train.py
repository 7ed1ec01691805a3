"""GPU parity of the whole training step (forward in train mode + loss + hand-written backward + BN statistics + AdamW)
through the reference's own entry points (TDEEDModel.epoch / get_optimizer) against

  * tests/golden/train_*.npz — loss, gradients and BN buffers produced by the UNMODIFIED reference (CPU fp32), and
  * the oracle (oracle/train_oracle.py) evaluated on the GPU box in float64, which gives the noise floor: the reference's own
    fp32 gradients differ from the float64 ones by 1-3e-2 relative-to-max on these nets (a ReLU / max-pool routing flip
    moves whole rows of the gradient), so the end-to-end bound on the fp32 kernels is  E <= 4 * max_t E_ref(t) + 1e-4
    per tensor, while loss / logits / BN statistics are held to 1e-4 and every kernel to 2e-4 in test_gpu_train_ops.py.
"""
import os
import random
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import tdeed_oracle as O
import train_oracle as TO
from gen_golden_train import TRAIN_CASES
from test_oracle_golden import train_case_from_golden, rel_err


def _args(cfg):
    return Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=cfg.radi_displacement,
                     feature_arch=cfg.feature_arch, clip_len=cfg.clip_len, n_layers=cfg.n_layers, sgp_ks=cfg.sgp_ks,
                     sgp_r=cfg.sgp_r, num_classes=cfg.num_classes, crop_dim=cfg.crop_dim)


def _model(cfg, sd):
    import contextlib
    import io
    from model.model import TDEEDModel
    with contextlib.redirect_stdout(io.StringIO()):
        m = TDEEDModel(device='cuda', args=_args(cfg))
    m.load(sd)
    m._model.augmentation = torch.nn.Identity()
    for mod in m._model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def assert_grad_parity(e_mine, e_ref, tag=''):
    """Per-tensor relative-to-max gradient errors against the float64 oracle: kernels (e_mine) vs the reference arithmetic in
    fp32 (e_ref, torch on the same GPU).  A ReLU / max-pool / argmax routing decision that lands on the other side in fp32
    changes isolated gradient entries by O(1) — the reference's own fp32 run shows 1e-2 .. 8e-1 on these nets, and differently
    from run to run (its kernels are not deterministic) — so the bound is on the bulk of the distribution, with a loose cap on
    the worst tensor:  median <= 3x reference median,  90th percentile <= max(4x reference's, 2e-3),  worst <= max(4x ref, 0.5)."""
    a, r = np.sort(list(e_mine.values())), np.sort(list(e_ref.values()))
    med_a, med_r = float(np.median(a)), float(np.median(r))
    p90_a, p90_r = float(a[int(0.9 * (len(a) - 1))]), float(r[int(0.9 * (len(r) - 1))])
    worst = max(e_mine, key=e_mine.get)
    print('%s: kernels median %.2e p90 %.2e worst %.2e (%s) | fp32 reference arithmetic median %.2e p90 %.2e worst %.2e'
          % (tag, med_a, p90_a, e_mine[worst], worst, med_r, p90_r, float(r[-1])))
    assert med_a <= 3 * med_r + 1e-5, (med_a, med_r)
    assert p90_a <= max(4 * p90_r, 2e-3), (p90_a, p90_r)
    assert e_mine[worst] <= max(4 * float(r[-1]), 0.5), (worst, e_mine[worst])


def _oracle_grads(sd, cfg, frames, label, labelD, dtype):
    dev = 'cuda'
    s = {k: (v.to(dev).to(dtype) if v.dtype.is_floating_point else v.to(dev)) for k, v in sd.items()}
    lab = label.to(dev)
    if lab.dtype.is_floating_point:
        lab = lab.to(dtype)
    loss, logits, displ, grads, after = TO.train_forward_backward(
        s, cfg, frames.to(dev).to(dtype), lab, labelD.to(dev).to(dtype) if labelD is not None else None)
    return loss, logits, grads, after


@pytest.mark.parametrize('name', sorted(TRAIN_CASES))
def test_train_step_fp32_matches_reference(name, golden_dir):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg, sd, frames, label, labelD, g = train_case_from_golden(name, golden_dir)
    m = _model(cfg, sd)
    m._model.train()
    lab = label.cuda()
    lab = lab.reshape(-1) if lab.dim() == 2 else lab.reshape(-1, lab.shape[-1])
    loss = m._model.train_step(frames.cuda(), lab, labelD.cuda() if labelD is not None else None, fg_weight=5, precision='fp32')
    torch.cuda.synchronize()
    # loss and BN buffers: tight
    assert abs(float(loss[0]) - float(g['loss'])) < 1e-4 * max(1.0, abs(float(g['loss'])))
    after = m.state_dict()
    for k in g.files:
        if k.startswith('after/'):
            np.testing.assert_allclose(after[k[6:]].cpu().numpy(), g[k], rtol=2e-4, atol=2e-5)
    # gradients: noise floor from the float64 oracle
    _, logits64, g64, _ = _oracle_grads(sd, cfg, frames, label, labelD, torch.float64)
    _, _, g32, _ = _oracle_grads(sd, cfg, frames, label, labelD, torch.float32)
    mine = {n: p.grad for n, p in m._model.named_parameters()}
    assert rel_err(m._model._last_train[0].double().cpu().numpy(), logits64.cpu().numpy()) < 1e-4
    e_ref = {n: rel_err(g32[n].double().cpu().numpy(), g64[n].cpu().numpy()) for n in g64 if g64[n].numel() >= 16}
    e_mine = {n: rel_err(mine[n].double().cpu().numpy(), g64[n].cpu().numpy()) for n in g64 if g64[n].numel() >= 16}
    assert_grad_parity(e_mine, e_ref, name)
    # and directly against the reference's own gradients (golden, CPU fp32): L2 norm of every parameter tensor — 90 % of the
    # tensors within 1 %, all within 50 % (same routing-flip caveat)
    names = [str(n) for n in g['grad_names']]
    dev_l2 = []
    for n, (s_, sa, l2) in zip(names, g['grad_stats']):
        a = mine[n].double().cpu().numpy()
        if a.size >= 16:
            dev_l2.append(abs(np.sqrt((a * a).sum()) - l2) / max(l2, 1e-6))
    dev_l2 = np.sort(dev_l2)
    assert dev_l2[int(0.9 * (len(dev_l2) - 1))] < 1e-2 and dev_l2[-1] < 0.5, (dev_l2[int(0.9 * (len(dev_l2) - 1))], dev_l2[-1])


def _cosines(grads, g64):
    cos = {}
    for n, b in g64.items():
        a, b = grads[n].double().reshape(-1), b.reshape(-1)
        if a.numel() >= 64:
            cos[n] = float((a @ b) / (a.norm() * b.norm() + 1e-30))
    return cos


def test_train_step_bf16_vs_reference_autocast(golden_dir):
    """bf16 storage + tensor-core GEMMs.  Stated tolerance (relative to the reference's own autocast error):
    the gradient direction of every tensor (cosine against the float64 oracle) must be at least as good as what the
    reference modules give under torch.autocast(bfloat16) on the same inputs, minus 0.1 / 0.05 (min / median) slack; the
    loss within 2x the autocast run's own loss error (and 6 %)."""
    cfg, sd, frames, label, labelD, g = train_case_from_golden('rny002_gsf_displ', golden_dir)
    m = _model(cfg, sd)
    m._model.train()
    lab = label.cuda().reshape(-1)
    loss = m._model.train_step(frames.cuda(), lab, labelD.cuda(), fg_weight=5, precision='bf16')
    mine = {n: p.grad for n, p in m._model.named_parameters()}
    loss64, _, g64, _ = _oracle_grads(sd, cfg, frames, label, labelD, torch.float64)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        loss_ac, _, g_ac, _ = _oracle_grads(sd, cfg, frames, label, labelD, torch.float32)
    c_mine, c_ref = _cosines(mine, g64), _cosines(g_ac, g64)
    print('loss: kernels bf16 %.4f | autocast reference %.4f | f64 %.4f' % (float(loss[0]), loss_ac, loss64))
    print('gradient cosine vs f64: kernels min %.3f median %.3f | autocast reference min %.3f median %.3f'
          % (min(c_mine.values()), float(np.median(list(c_mine.values()))), min(c_ref.values()), float(np.median(list(c_ref.values())))))
    assert abs(float(loss[0]) - loss64) < max(2 * abs(loss_ac - loss64), 6e-2 * abs(loss64))
    assert float(np.median(list(c_mine.values()))) >= float(np.median(list(c_ref.values()))) - 0.05
    assert min(c_mine.values()) >= min(c_ref.values()) - 0.1


def test_epoch_mixup_glue_matches_reference_loss(golden_dir):
    """TDEEDModel.epoch(optimizer=...) end to end: mixup of frames / labels (model/model.py:228-254) + train_step."""
    name = 'rny002_gsf_mixup'
    kw, _, wseed, iseed, _ = TRAIN_CASES[name]
    g = np.load(os.path.join(golden_dir, 'train_%s.npz' % name))
    cfg = O.Config(**kw)
    m = _model(cfg, O.random_state(cfg, wseed))
    m.train_precision = 'fp32'
    batch = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('batch_')}

    class Recording:
        def zero_grad(self):
            pass

        def step(self):
            pass

    random.seed(iseed)
    loss = m.epoch([batch], optimizer=Recording(), scaler=None, fg_weight=5)
    assert abs(loss - float(g['loss'])) < 1e-4 * abs(float(g['loss']))


def test_epoch_trains_with_fused_adamw_and_scheduler():
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=8, n_layers=2, sgp_ks=5, sgp_r=4, num_classes=4, radi_displacement=2,
                   crop_dim=64)
    m = _model(cfg, O.random_state(cfg, 5))
    keys_before = list(m.state_dict().keys())
    gen = torch.Generator().manual_seed(0)
    batch = {'frame': torch.randint(0, 256, (2, 8, 3, 64, 80), generator=gen, dtype=torch.uint8),
             'label': torch.randint(0, 5, (2, 8), generator=gen), 'labelD': torch.randint(-2, 3, (2, 8), generator=gen)}
    opt, scaler = m.get_optimizer({'lr': 3e-4})
    assert isinstance(opt, torch.optim.Optimizer)
    sched = torch.optim.lr_scheduler.ChainedScheduler([torch.optim.lr_scheduler.LinearLR(opt, start_factor=0.1, total_iters=3),
                                                       torch.optim.lr_scheduler.CosineAnnealingLR(opt, 40)])
    w0 = m.state_dict()['_temp_fine._sgp.0.mlp.0.weight'].clone()
    losses = [m.epoch([batch, batch], optimizer=opt, scaler=scaler, lr_scheduler=sched, acc_grad_iter=1) for _ in range(12)]
    print('losses', ['%.3f' % v for v in losses])
    assert all(np.isfinite(losses)) and min(losses[-3:]) < 0.8 * losses[0]
    assert list(m.state_dict().keys()) == keys_before
    assert not torch.equal(m.state_dict()['_temp_fine._sgp.0.mlp.0.weight'], w0)
    assert int(m.state_dict()['_features.stem.bn.num_batches_tracked']) == 24
    # gradient accumulation == one step on the concatenated batch statistics-wise is not identical (BN), just check it runs
    l2 = m.epoch([batch, batch], optimizer=opt, scaler=scaler, acc_grad_iter=2)
    assert np.isfinite(l2)
    # inference sees the trained weights (engine cache invalidated by the fused optimizer)
    cls, probs = m.predict(batch['frame'][:1], use_amp=False)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ocls, oprobs = O.predict(sd, cfg, batch['frame'][:1])
    assert np.abs(probs - oprobs).max() < 1e-3
    # evaluation epoch still works after training
    assert np.isfinite(m.epoch([batch], optimizer=None))


def test_train_step_cuda_graph_equals_eager(golden_dir):
    """The captured forward+loss+backward graph (3rd call of a signature on) writes bit-identical gradients."""
    cfg, sd, frames, label, labelD, g = train_case_from_golden('rny002_gsf_displ', golden_dir)
    m = _model(cfg, sd)
    m._model.train()
    lab, fr, ld = label.cuda().reshape(-1), frames.cuda(), labelD.cuda()
    flat = m._model.flat_params()
    buffers0 = {k: v.clone() for k, v in m._model.named_buffers()}

    def reset_buffers():
        with torch.no_grad():
            for k, v in m._model.named_buffers():
                v.copy_(buffers0[k])

    loss_e = m._model.train_step(fr, lab, ld, precision='bf16', use_graph=False).clone()
    g_e = flat.g.clone()
    reset_buffers()
    for i in range(3):                                   # warm (eager), capture + replay, replay
        loss_g = m._model.train_step(fr, lab, ld, precision='bf16', use_graph=True).clone()
        assert torch.equal(flat.g, g_e), i
        assert torch.equal(loss_g, loss_e)
        if i < 2:
            reset_buffers()
    ent = [v for v in m._model._train_graphs.values() if isinstance(v, dict)]
    assert len(ent) == 1 and 'graph' in ent[0]
    # running statistics advanced exactly once since the last reset
    assert int(m.state_dict()['_features.stem.bn.num_batches_tracked']) == int(buffers0['_features.stem.bn.num_batches_tracked']) + 1


@pytest.mark.parametrize('soft', [False, True])
def test_double_head_joint_training_step(soft):
    """Joint-dataset training with the double head (model/model.py:278-306, train_tdeed.py:145-148): per-sample head selection
    by batch['dataset'], update_labels_2heads label shift, FC2Layers with one dropout per head.  Checked against autograd of the
    oracle network + the reference's loss loop restated in torch, in float64."""
    import torch.nn.functional as F
    n1, n2 = 5, 7
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=8, n_layers=2, sgp_ks=5, sgp_r=4, num_classes=n1 - 1, radi_displacement=2,
                   crop_dim=None, double_head=[n1, n2])
    sd = O.random_state(cfg, 9)
    import contextlib
    import io
    from model.model import TDEEDModel, update_labels_2heads
    args = _args(cfg)
    with contextlib.redirect_stdout(io.StringIO()):
        m = TDEEDModel(device='cuda', args=args)
    m._model.update_pred_head([n1, n2])
    m._num_classes = n1 + n2
    m.load(sd)
    m._model.augmentation = torch.nn.Identity()
    for mod in m._model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m._model.train()
    gen = torch.Generator().manual_seed(3)
    B, T_ = 3, 8
    frames = torch.randint(0, 256, (B, T_, 3, 64, 64), generator=gen, dtype=torch.uint8)
    dataset = [1, 2, 2]
    label = torch.stack([torch.randint(0, n1 if d == 1 else n2, (T_,), generator=gen) for d in dataset])
    labelD = torch.randint(-2, 3, (B, T_), generator=gen).float()
    label = update_labels_2heads(label.clone(), dataset, n1 - 1)
    K = n1 + n2
    if soft:
        onehot = F.one_hot(label, K).float()
        other = F.one_hot(update_labels_2heads(torch.stack([torch.randint(0, n1 if d == 1 else n2, (T_,), generator=gen)
                                                             for d in dataset]), dataset, n1 - 1), K).float()
        target = 0.7 * onehot + 0.3 * other
        lab_in = target.reshape(-1, K).cuda()
    else:
        target = label
        lab_in = label.reshape(-1).cuda()
    loss = m._model.train_step(frames.cuda(), lab_in, labelD.cuda(), fg_weight=5, precision='fp32', dataset=dataset)
    # float64 reference: oracle network + the reference's per-sample loss loop
    s64 = {k: (v.cuda().double().requires_grad_(True) if v.dtype.is_floating_point and 'running' not in k else v.cuda())
           for k, v in sd.items()}
    logits, displ = O.forward(s64, cfg, frames.cuda().double(), train=True)
    w = torch.tensor([1.] + [5.] * (K - 1), dtype=torch.float64, device='cuda')
    ref = 0.
    for i in range(B):
        if dataset[i] == 1:
            tgt = target[i][:, :n1] if soft else target[i]
            ref = ref + F.cross_entropy(logits[i][:, :n1], tgt.cuda().double() if soft else tgt.cuda(), weight=w[:n1]) / B
        else:
            tgt = target[i][:, n1:] if soft else target[i] - n1
            ref = ref + F.cross_entropy(logits[i][:, n1:], tgt.cuda().double() if soft else tgt.cuda(), weight=w[:n2]) / B
    ref = ref + F.mse_loss(displ, labelD.cuda().double(), reduction='none').mean()
    ref.backward()
    assert abs(float(loss[0]) - float(ref)) < 1e-4 * max(1.0, abs(float(ref)))
    for name in ('_pred_fine._fc1._fc_out.weight', '_pred_fine._fc2._fc_out.weight', '_pred_fine._fc2._fc_out.bias',
                 '_pred_displ._fc_out.weight', '_temp_fine._sgp.4.mlp.2.weight', '_temp_fine._sgpMixer.0.concat_fc.weight'):
        got = dict(m._model.named_parameters())[name].grad
        assert rel_err(got.double().cpu().numpy(), s64[name].grad.cpu().numpy()) < 2e-3, name
    # and through epoch() with the reference's batch schema
    batch = {'frame': frames, 'label': update_labels_2heads(label.clone(), dataset, -1 - n1) if False else
             torch.stack([label[i] - (n1 if dataset[i] == 2 else 0) for i in range(B)]), 'labelD': labelD.long(), 'dataset': dataset}
    m.train_precision = 'fp32'
    m._args.num_classes = n1 - 1

    class Recording:
        def zero_grad(self):
            pass

        def step(self):
            pass

    if not soft:
        l2 = m.epoch([batch], optimizer=Recording(), scaler=None, fg_weight=5)
        assert abs(l2 - float(ref)) < 2e-3 * max(1.0, abs(float(ref)))      # (BN running stats moved once in between: loss identical)


def _variant_case(arch, shape, radi, seed):
    b, t, h, w = shape
    cfg = O.Config(feature_arch=arch, clip_len=t, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=3, radi_displacement=radi, crop_dim=None)
    sd = O.random_state(cfg, seed)
    gen = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (b, t, 3, h, w), generator=gen, dtype=torch.uint8).float()
    label = torch.randint(0, 4, (b, t), generator=gen)
    labelD = torch.randint(-radi, radi + 1, (b, t), generator=gen).float() if radi else None
    m = _model(cfg, sd)
    m._model.train()
    loss = m._model.train_step(frames.cuda(), label.cuda().reshape(-1), labelD.cuda() if radi else None, fg_weight=5, precision='fp32')
    l64, logits64, g64, _ = _oracle_grads(sd, cfg, frames, label, labelD, torch.float64)
    _, _, g32, _ = _oracle_grads(sd, cfg, frames, label, labelD, torch.float32)
    assert abs(float(loss[0]) - l64) < 1e-4 * max(1.0, abs(l64))
    mine = {n: p.grad for n, p in m._model.named_parameters()}
    e_ref = {n: rel_err(g32[n].double().cpu().numpy(), g64[n].cpu().numpy()) for n in g64 if g64[n].numel() >= 16}
    e_mine = {n: rel_err(mine[n].double().cpu().numpy(), g64[n].cpu().numpy()) for n in g64 if g64[n].numel() >= 16}
    return cfg, sd, frames, label, labelD, l64, e_mine, e_ref


@pytest.mark.parametrize('arch,shape,radi,seed', [('rny002_gsm', (2, 10, 64, 64), 0, 18), ('rny002_gsf', (2, 6, 52, 76), 2, 13),
                                                  ('rny008_gsf', (1, 7, 45, 33), 1, 13)])
def test_train_step_variants_vs_f64_oracle(arch, shape, radi, seed):
    """Variants without a reference golden — GSM (the reference's _GSM needs CUDA tensors to even run), odd frame sizes (every
    stride-2 stage sees odd extents: parity views / reflect-free padding paths), 800MF with odd sizes — against autograd of the
    float64 oracle with the same noise-floor rule as the golden cases."""
    cfg, sd, frames, label, labelD, l64, e_mine, e_ref = _variant_case(arch, shape, radi, seed)
    assert_grad_parity(e_mine, e_ref, '%s %s' % (arch, shape))
    # On these tiny random nets most inputs contain a ReLU / max-pool routing decision that sits within fp32 round-off of its
    # threshold; it flips between ANY two fp32 evaluation orders (the reference arithmetic included: see the fp32-oracle column
    # printed by assert_grad_parity) and moves isolated gradient tensors by 1e-2 .. 2e-1.  Which seeds are flip-free therefore
    # depends on the summation order of every kernel, so no seed is hard-wired: the kernels are deterministic, hence on SOME
    # seed of a small fixed pool EVERY gradient tensor must agree with the float64 oracle to 1e-3 — a systematic error in any
    # backward kernel would break this on all of them.
    worst = {seed: max(e_mine.values())}
    for s2 in (11, 12, 13, 14, 15, 16, 17, 18):
        if min(worst.values()) < 1e-3:
            break
        if s2 not in worst:
            worst[s2] = max(_variant_case(arch, shape, radi, s2)[6].values())
    assert min(worst.values()) < 1e-3, 'no flip-free seed: worst per-tensor error by seed %s' % worst
    # bf16 path on the same (odd) geometry: runs, finite, loss close
    m2 = _model(cfg, sd)
    m2._model.train()
    lb = m2._model.train_step(frames.cuda(), label.cuda().reshape(-1), labelD.cuda() if labelD is not None else None, fg_weight=5,
                              precision='bf16')
    assert abs(float(lb[0]) - l64) < 0.1 * abs(l64)
    assert all(torch.isfinite(p.grad).all() for p in m2._model.parameters())
