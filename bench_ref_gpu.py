#!/usr/bin/env python3
"""`ref_gpu` leg of bench.py: the reference network executed by STOCK torch on the
same B200 — cuDNN / cuBLAS kernels, fp16 autocast as the reference does (model/model.py:343-346) and bf16 autocast,
channels_last activations.  This is "the existing Blackwell library path" every hand-written kernel family has to beat.

/root/reference does not exist on the GPU box, so the network is the oracle port (oracle/tdeed_oracle.py — the functional
restatement pinned to the unmodified reference by tests/golden/*), moved to the GPU.  It is a BASELINE leg: nothing here is
on the product path.  Per-family times come from CUDA events around the oracle's own building blocks (monkeypatched
here, not in oracle/): conv1x1 (+BN+ReLU, as separate library kernels), grouped 3x3, stem, gate-shift, SE (bottleneck
time minus its convs / gate-shift), SGP temporal stack, heads.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


class _FamilyTimer:
    def __init__(self):
        self.on = False
        self.events = []

    def wrap(self, label_fn, fn):
        def inner(*a, **k):
            if not self.on:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            self.events.append((label_fn(*a, **k), e0, e1))
            return out
        return inner

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for label, e0, e1 in self.events:
            out[label] = out.get(label, 0.0) + e0.elapsed_time(e1)
        self.events = []
        return out


def _install(O, timer):
    """Time the oracle's building blocks; returns an undo function."""
    orig = dict(_cna=O._cna, gate_shift=O.gate_shift, bottleneck=O.bottleneck, ed_sgp_mixer=O.ed_sgp_mixer, heads=O.heads,
                preprocess=O.preprocess)

    def cna_label(x, sd, p, stride=1, groups=1, act=True, train=False):
        if p.endswith('.stem'):
            return 'stem'
        if groups > 1:
            return 'conv3x3g'
        return 'conv1x1_ds' if p.endswith('.downsample') else 'conv1x1'
    O._cna = timer.wrap(cna_label, orig['_cna'])
    O.gate_shift = timer.wrap(lambda *a, **k: 'gsf', orig['gate_shift'])
    O.bottleneck = timer.wrap(lambda *a, **k: '_bottleneck', orig['bottleneck'])
    O.ed_sgp_mixer = timer.wrap(lambda *a, **k: 'sgp', orig['ed_sgp_mixer'])
    O.heads = timer.wrap(lambda *a, **k: 'heads', orig['heads'])

    def preprocess_cl(frames, cfg, flip=False, crop_offsets=None):
        return orig['preprocess'](frames, cfg, flip, crop_offsets).contiguous(memory_format=torch.channels_last)
    O.preprocess = timer.wrap(lambda *a, **k: 'preprocess', preprocess_cl)

    def undo():
        for k, v in orig.items():
            setattr(O, k, v)
    return undo


def _families(raw):
    fam = {k: v for k, v in raw.items() if not k.startswith('_')}
    inner = sum(raw.get(k, 0.0) for k in ('conv1x1', 'conv1x1_ds', 'conv3x3g', 'gsf'))
    fam['se_and_residual'] = max(0.0, raw.get('_bottleneck', 0.0) - inner)
    return {k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}


def run_inference(config_name, state, frame_hw, dev, clips=8, reps=3, dtypes=(torch.float16, torch.bfloat16)):
    """clips/s of the oracle forward + softmax/scatter on the GPU under autocast, plus per-family ms per batch."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import tdeed_oracle as O
    cfg = O.named_config(config_name)
    sd = {k: v.detach().to(dev) for k, v in state.items()}
    g = torch.Generator(device=dev).manual_seed(0)
    frames = torch.randint(0, 256, (clips, 100, 3) + tuple(frame_hw), generator=g, dtype=torch.uint8, device=dev)
    torch.backends.cudnn.benchmark = True
    timer = _FamilyTimer()
    undo = _install(O, timer)
    out = {'clips_per_batch': clips, 'reps': reps, 'memory_format': 'channels_last',
           'what': 'oracle port of the reference network on stock torch %s (cuDNN/cuBLAS), autocast' % torch.__version__}
    try:
        for dt in dtypes:
            name = str(dt).split('.')[-1]

            def step():
                with torch.no_grad(), torch.autocast('cuda', dtype=dt):
                    logits, displ = O.forward(sd, cfg, frames)
                    p = torch.softmax(logits.float(), dim=2)
                    if displ is not None:      # process_prediction as one scatter (the reference's Python loop would dominate)
                        t = p.shape[1]
                        idx = (torch.arange(t, device=dev)[None] - torch.round(displ.float()).long()).clamp(0, t - 1)
                        aux = torch.zeros_like(p)
                        aux.scatter_reduce_(1, idx[:, :, None].expand_as(p), p, reduce='amax', include_self=True)
                        p = aux
                return p
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            timer.on = True
            step()
            fam = _families(timer.totals())
            timer.on = False
            out[name] = {'clips_per_s': clips / ms * 1e3, 'ms_per_batch': ms, 'families_ms_per_batch': fam}
    finally:
        undo()
        torch.backends.cudnn.benchmark = False
    return out


def run_train(config_name, state, dev, clips=8, reps=3, dtype=torch.bfloat16):
    """clips/s of forward (train-mode BN) + weighted CE + autograd backward + torch.optim.AdamW(fused) under autocast."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import tdeed_oracle as O
    import torch.nn.functional as F
    cfg = O.named_config(config_name)
    work, params = {}, []
    for k, v in state.items():
        t = v.detach().to(dev).clone()
        if t.dtype.is_floating_point and not k.endswith(('running_mean', 'running_var')):
            t.requires_grad_(True)
            params.append(t)
        work[k] = t
    opt = torch.optim.AdamW(params, lr=1e-4, fused=True)
    g = torch.Generator(device=dev).manual_seed(0)
    frames = torch.randint(0, 256, (clips, 100, 3, 224, 224), generator=g, dtype=torch.uint8, device=dev)
    kk = cfg.num_classes + 1
    soft = torch.softmax(torch.randn((clips * 100, kk), generator=g, device=dev) * 3, dim=1)
    weight = torch.tensor([1.] + [5.] * (kk - 1), device=dev)
    undo = _install(O, _FamilyTimer())         # only for the channels_last preprocess

    def step():
        with torch.autocast('cuda', dtype=dtype):
            logits, _ = O.forward(work, cfg, frames, train='update')
            loss = F.cross_entropy(logits.reshape(-1, kk).float(), soft, weight=weight)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    try:
        torch.backends.cudnn.benchmark = True
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    finally:
        undo()
        torch.backends.cudnn.benchmark = False
    return {'clips_per_s': clips / ms * 1e3, 'ms_per_step': ms, 'clips_per_step': clips, 'dtype': str(dtype).split('.')[-1],
            'what': 'oracle port: train-mode forward + weighted soft-label CE + autograd backward + fused torch AdamW; no augmentation / mixup'}
