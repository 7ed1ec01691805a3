#!/usr/bin/env python3
"""Training-step benchmark (BASELINE.json configs[2]: FineGym_big — RegNetY-800MF + GSF, bf16 training step,
data-parallel 1/2/4/8 B200).  `python bench.py --workload train ...` dispatches here; same launch contract and JSON line
as bench.py (one process per GPU under torchrun, barrier + synchronize around the timed region, max over ranks).

One step = TDEEDModel.epoch's body for one batch (model/model.py:215-324 of the reference): mixup of the batch with its
partner clips, random crop, per-clip augmentation, forward in train mode, weighted CE on soft labels, backward, NCCL
all-reduce of the flat gradient (N > 1), fused AdamW.  8 clips of 100x3x224x224 uint8 per GPU (weak scaling).
`value` = clips/s with the batch resident in HBM; `e2e` = the same through TDEEDModel.epoch with pinned host batches
(H2D of frame + frame2 + labels inside the timed region, loss read back every step).
"""
import contextlib
import io
import json
import os
import random
import sys
import time
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))

TRAIN_CONFIG = dict(name='FineGym_big', feature_arch='rny008_gsf', clip_len=100, n_layers=3, sgp_ks=9, sgp_r=4, num_classes=32,
                    radi_displacement=0, crop_dim=224)
CLIPS_PER_GPU = 8
# kernel launches behind one C-ABI call (default 1); used for the gpu_launches claim
LAUNCHES = dict(tdeed_bn_stats=2, tdeed_bn_act_bwd=3, tdeed_gemm_tn=2, tdeed_colsum=2, tdeed_se_train_fwd=3, tdeed_se_bwd=3,
                tdeed_gsf_cat_fwd=4, tdeed_gsf_bwd=12, tdeed_conv3x3g_bwd_weight=2, tdeed_stem_bwd_weight=2, tdeed_chan_ln_bwd=2,
                tdeed_sgp_branch_bwd=4, tdeed_groupnorm_bwd=2, tdeed_pool_posenc_bwd=2)


def train_args():
    c = TRAIN_CONFIG
    return Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=c['radi_displacement'],
                     feature_arch=c['feature_arch'], clip_len=c['clip_len'], n_layers=c['n_layers'], sgp_ks=c['sgp_ks'],
                     sgp_r=c['sgp_r'], num_classes=c['num_classes'], crop_dim=c['crop_dim'])


class _Profiler:
    """Wraps the module-level wrappers of tdeed_b200.ops / train_ops with CUDA-event timers and launch counters."""

    def __init__(self):
        from tdeed_b200 import ops, train_ops, _lib
        self.mods = (ops, train_ops)
        self.records = None
        self.calls = 0
        self.launches = 0
        self._orig = {}
        lib = _lib.load()
        prof = self

        class CountingLib:
            def __getattr__(self, name):
                fn = getattr(lib, name)
                if not name.startswith('tdeed_') or name.endswith(('_floats', '_bytes', 'last_error', 'abi_version')):
                    return fn

                def call(*a):
                    prof.launches += LAUNCHES.get(name, 1)
                    return fn(*a)
                return call

        self._lib_mod, self._lib_load, self._proxy = _lib, _lib.load, CountingLib()

    def install(self):
        self._lib_mod.load = lambda: self._proxy
        for mod in self.mods:
            for name, fn in list(vars(mod).items()):
                if callable(fn) and not name.startswith('_') and getattr(fn, '__module__', None) == mod.__name__:
                    self._orig[(mod, name)] = fn
                    setattr(mod, name, self._wrap(name, fn))

    def remove(self):
        self._lib_mod.load = self._lib_load
        for (mod, name), fn in self._orig.items():
            setattr(mod, name, fn)

    def _wrap(self, name, fn):
        def wrapped(*a, **k):
            if self.records is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            self.records.append((name, e0, e1))
            return out
        return wrapped


def make_batch(rank, dev=None, pinned=False):
    g = torch.Generator().manual_seed(4321 + rank)
    b, t, k = CLIPS_PER_GPU, TRAIN_CONFIG['clip_len'], TRAIN_CONFIG['num_classes'] + 1

    def labels():
        lab = torch.zeros((b, t), dtype=torch.int64)
        hit = torch.rand((b, t), generator=g) < 0.016            # ~16 events per 1000 frames
        lab[hit] = torch.randint(1, k, (int(hit.sum()),), generator=g)
        return lab

    batch = {'frame': torch.randint(0, 256, (b, t, 3, 224, 224), generator=g, dtype=torch.uint8), 'label': labels(),
             'frame2': torch.randint(0, 256, (b, t, 3, 224, 224), generator=g, dtype=torch.uint8), 'label2': labels()}
    if pinned:
        batch = {k_: v.pin_memory() for k_, v in batch.items()}
    if dev is not None:
        batch = {k_: v.to(dev) for k_, v in batch.items()}
    return batch


def run_train_reference(args):
    """CPU oracle training step (train-mode forward + loss + autograd backward), rank 0 only, bounded sample: 1 clip."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import tdeed_oracle as O
    import train_oracle as TO
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = O.named_config(TRAIN_CONFIG['name'])
    sd = O.random_state(cfg, 0)
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (1, 100, 3, 224, 224), generator=g, dtype=torch.uint8).float()
    label = torch.randint(0, cfg.num_classes + 1, (1, 100), generator=g)
    for _ in range(args.warmup):
        TO.train_forward_backward(sd, cfg, frames, label, None)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        TO.train_forward_backward(sd, cfg, frames, label, None)
    dt = (time.perf_counter() - t0) / args.steps
    value = 1.0 / dt
    sample = '%d steps x 1 clip (100x3x224x224) forward(train)+loss+autograd backward of the CPU oracle (no optimizer step)' % args.steps
    print(json.dumps({
        'impl': 'reference', 'metric': 'train_clips_per_s', 'value': value, 'unit': 'clips/s', 'frames_per_s': value * 100,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': TRAIN_CONFIG['name'] + ' training step', 'clips_per_step': 1},
        'cpu_baseline': {'value': value, 'unit': 'clips/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def run_train(args, quiet=False):
    """Returns the JSON dict on rank 0 (None elsewhere)."""
    from bench import ClockSampler, peaks, randomize_
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1 and os.environ.get('TDEED_NUMA_BIND', '1') != '0':
        from tdeed_b200.parallel import bind_to_gpu_numa
        bind_to_gpu_numa(local)                    # idempotent; before the pinned host batches are allocated
    dev = torch.device('cuda', local)
    import torch.distributed as dist
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)

    from model.model import TDEEDModel
    prof = _Profiler()
    prof.install()
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)             # the conv / linear weights come from torch's default init: the same on every run and rank
        model = TDEEDModel(device='cuda:%d' % local, args=train_args())
    randomize_(model._model, seed=0)                      # same weights on every rank (DP replicas)
    model.train_precision = args.precision
    if args.no_aug:
        model._model.augmentation = torch.nn.Identity()
    opt, scaler = model.get_optimizer({'lr': 1e-4})
    dev_batch = make_batch(rank, dev=dev)
    host_batch = make_batch(rank, pinned=True)
    random.seed(1 + rank)
    torch.manual_seed(1 + rank)

    class _Loader(list):           # tqdm(loader) wants len()
        pass

    def step(batch, n=1):
        """n training steps = one TDEEDModel.epoch over n batches (the loss is read back once, at the end of the epoch)."""
        with contextlib.redirect_stderr(io.StringIO()):       # tqdm bar
            return model.epoch(_Loader([batch] * n), optimizer=opt, scaler=scaler, fg_weight=5)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    if not args.no_aug:
        # untimed: run every transform of the augmentation pipeline once (each fires with p = 0.25 per clip, so a short warm-up
        # may never reach e.g. GaussianBlur, whose first call pays cuDNN's one-time initialisation)
        dummy = torch.rand((100, 3, 224, 224), device=dev)
        for tr in model._model.augmentation.transforms:
            inner = getattr(tr, 'transforms', None)
            for t_ in (inner if inner is not None else [tr]):
                t_(dummy)
        del dummy
    loss = step(dev_batch, args.warmup)
    barrier()
    if getattr(args, 'ncu_step', False):
        model._model.use_train_graph = False
        step(dev_batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(dev_batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    loss = step(dev_batch, args.steps)
    e1.record()
    barrier()
    dev_s = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    clocks = sampler.stop() if rank == 0 else None

    step(host_batch)
    barrier()
    t0 = time.perf_counter()
    step(host_batch, args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    families, roofline = {}, None
    # the timed steps replay a CUDA graph (one host call for forward+loss+backward); to attribute time to kernel families and to
    # count the kernels inside the graph, one more step is run eagerly with CUDA-event brackets around every C-ABI call
    model._model.use_train_graph = False
    step(dev_batch)
    barrier()
    l0 = prof.launches
    if rank == 0:
        prof.records = []
    step(dev_batch)
    torch.cuda.synchronize()
    launches = (prof.launches - l0) * args.steps
    if rank == 0:
        for name, a, b in prof.records:
            f = families.setdefault(name, dict(ms=0.0, calls=0))
            f['ms'] += a.elapsed_time(b)
            f['calls'] += 1
        prof.records = None
        # dominant family: the weight-gradient GEMMs; algorithmic FLOPs = forward 1x1-conv / linear FLOPs (SURVEY 8d:
        # 126.05 GFLOP 1x1 + 6.38 SGP per clip)
        pk = peaks()
        top = max(families, key=lambda k_: families[k_]['ms'])
        total_ms = sum(v['ms'] for v in families.values())
        sec = families[top]['ms'] / 1e3
        # algorithmic work of the families that can dominate (FineGym_big, per step of CLIPS_PER_GPU clips):
        #   conv outputs of the backbone: 5.24 M elements per frame (every BatchNorm'd tensor)
        elems = 5.24e6 * 100 * CLIPS_PER_GPU
        # COMPULSORY bytes: bn_act_bwd reads dy and y once and writes dx (+ the 0.3 share of tensors that also emit a
        # residual gradient); the implementation's second pass over dy / y is traffic, not algorithm
        alg_bytes = {'bn_act_bwd': elems * 2 * (3 + 0.3), 'bn_act_fwd': elems * 2 * 2.3, 'bn_stats': elems * 2}
        alg_flops = {'gemm_tn': 2 * 0.5 * (126.05e9 + 6.38e9) * CLIPS_PER_GPU, 'gemm': 2 * (126.05e9 + 6.38e9) * CLIPS_PER_GPU}
        if top in alg_flops:
            roofline = {'kernel': top, 'bound': 'tensor', 'achieved': alg_flops[top] / sec / 1e12, 'peak': pk['tf'], 'unit': 'TFLOP/s'}
        else:
            roofline = {'kernel': top, 'bound': 'hbm', 'achieved': (alg_bytes[top] / sec / 1e9) if top in alg_bytes else None,
                        'peak': pk['hbm'], 'unit': 'GB/s'}
        roofline['frac'] = (roofline['achieved'] / roofline['peak']) if roofline['achieved'] else None
        roofline['traffic'] = None
        roofline['peak_source'] = pk['src']
        roofline['share_of_step'] = families[top]['ms'] / total_ms
    prof.remove()

    out = None
    if rank == 0:
        total = CLIPS_PER_GPU * args.steps * world
        value = total / dev_s
        frame_bytes = CLIPS_PER_GPU * 100 * 3 * 224 * 224
        out = {
            'metric': 'train_clips_per_s', 'value': value, 'unit': 'clips/s', 'frames_per_s': value * 100, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_s / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': TRAIN_CONFIG['name'] + ' training step', 'clips_per_gpu': CLIPS_PER_GPU,
                       'frame_shape': [100, 3, 224, 224], 'mixup': True, 'augmentation': (not args.no_aug) and 'fused kernels (train_aug.cu), reference distributions', 'optimizer': 'fused AdamW',
                       'l2_policy': 'activations of one step (>10 GB) larger than L2', 'parallelism': 'dp%d' % world},
            'e2e': {'value': total / e2e_s, 'unit': 'clips/s', 'h2d_bytes_per_step': int(2 * frame_bytes + 2 * CLIPS_PER_GPU * 100 * 8),
                    'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': None,
            'final_loss': float(loss),
            'kernel_families_ms_per_step': {k_: round(v['ms'], 3) for k_, v in sorted(families.items(), key=lambda kv: -kv[1]['ms'])},
        }
        if world == 1 and not getattr(args, 'no_ref_gpu', False):
            # stock-torch training step of the same network on the same GPU (cuDNN / cuBLAS + autograd), see bench_ref_gpu.py
            try:
                import bench_ref_gpu
                sd = {k_: v.detach() for k_, v in model.state_dict().items()}
                del dev_batch
                model._model._train_graphs.clear()
                torch.cuda.empty_cache()
                out['ref_gpu'] = {'train': bench_ref_gpu.run_train(TRAIN_CONFIG['name'], sd, dev, clips=CLIPS_PER_GPU)}
            except Exception as exc:
                out['ref_gpu'] = {'error': repr(exc)[:300]}
        if not quiet:
            print(json.dumps(out))
    return out
