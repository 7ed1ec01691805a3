"""Drop-in for the reference's model/model.py: same entry points (TDEEDModel, TDEEDModel.Impl,
update_labels_2heads), same constructor arguments, same state_dict layout — so the reference's
train_tdeed.py / evaluate_tdeed_challenge.py / util/eval.py run unchanged with this package first
on sys.path — while Impl.forward runs as hand-written sm_100a kernels (libtdeed_sm100.so through
tdeed_b200.InferenceEngine).  There is no PyTorch / CPU fallback: without the built library or a
CUDA device, forward raises.
"""
import os
import random
from contextlib import nullcontext

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from tqdm import tqdm

from model import regnet
from model.modules import (BaseRGBModel, EDSGPMIXERLayers, FCLayers, FC2Layers, step, process_prediction,
                           process_double_head, process_labels)
from model.shift import make_temporal_shift
from tdeed_b200.engine import EngineConfig, InferenceEngine


class TDEEDModel(BaseRGBModel):

    class Impl(nn.Module):

        def __init__(self, args=None):
            super().__init__()
            self._modality = args.modality
            assert self._modality == 'rgb', 'Only RGB supported for now'
            self._temp_arch = args.temporal_arch
            assert self._temp_arch in ['ed_sgp_mixer'], 'Only ed_sgp_mixer supported for now'
            self._radi_displacement = args.radi_displacement
            self._feature_arch = args.feature_arch
            assert 'rny' in self._feature_arch, 'Only rny supported for now'
            self._double_head = False
            self._args = args

            if self._feature_arch.startswith(('rny002', 'rny008')):
                features = regnet.create_model({
                    'rny002': 'regnety_002',
                    'rny008': 'regnety_008',
                }[self._feature_arch.rsplit('_', 1)[0]], pretrained=True)
                feat_dim = features.head.fc.in_features
                features.head.fc = nn.Identity()
                self._d = feat_dim
            else:
                raise NotImplementedError(args._feature_arch)

            self._require_clip_len = -1
            if self._feature_arch.endswith('_gsm'):
                make_temporal_shift(features, args.clip_len, mode='gsm')
                self._require_clip_len = args.clip_len
            elif self._feature_arch.endswith('_gsf'):
                make_temporal_shift(features, args.clip_len, mode='gsf')
                self._require_clip_len = args.clip_len

            self._features = features
            self._feat_dim = self._d
            self.temp_enc = nn.Parameter(torch.normal(mean=0, std=1 / args.clip_len, size=(args.clip_len, self._d)))
            if self._temp_arch == 'ed_sgp_mixer':
                self._temp_fine = EDSGPMIXERLayers(self._d, args.clip_len, num_layers=args.n_layers, ks=args.sgp_ks,
                                                   k=args.sgp_r, concat=True)
                self._pred_fine = FCLayers(self._feat_dim, args.num_classes + 1)
            else:
                raise NotImplementedError(self._temp_arch)
            if self._radi_displacement > 0:
                self._pred_displ = FCLayers(self._feat_dim, 1)

            self.croping = args.crop_dim
            # per-clip training augmentation of the reference (model/model.py:77-84); torchvision ops on device floats.
            # Assign nn.Identity() to disable (parity tests do).
            try:
                import torchvision.transforms as T
                self.augmentation = T.Compose([
                    T.RandomApply([T.ColorJitter(hue=0.2)], p=0.25),
                    T.RandomApply([T.ColorJitter(saturation=(0.7, 1.2))], p=0.25),
                    T.RandomApply([T.ColorJitter(brightness=(0.7, 1.2))], p=0.25),
                    T.RandomApply([T.ColorJitter(contrast=(0.7, 1.2))], p=0.25),
                    T.RandomApply([T.GaussianBlur(5)], p=0.25),
                    T.RandomHorizontalFlip(),
                ])
            except ImportError:      # pragma: no cover
                self.augmentation = nn.Identity()
            self._engines = {}
            self._engine_versions = {}
            self._flat = None
            self._train_engines = {}
            self._train_graphs = {}
            self.fused_augmentation = True   # False: run self.augmentation (torchvision ops) instead of train_aug.cu
            self.use_train_graph = True      # replay forward+loss+backward as a CUDA graph from the 3rd step of a signature on
            self.overlap_allreduce = True    # data parallel: all-reduce the temporal-stack gradients during the backbone backward
            self._grads_reduced = False
            self._train_calls = 0

        # ---- engine management -------------------------------------------------------------
        def engine_config(self):
            a = self._args
            dh = None
            if self._double_head:
                dh = [self._pred_fine._fc1._fc_out.out_features, self._pred_fine._fc2._fc_out.out_features]
            return EngineConfig(self._feature_arch, a.clip_len, a.n_layers, a.sgp_ks, a.sgp_r, a.num_classes,
                                self._radi_displacement, self.croping, double_head=dh)

        def _weights_version(self):
            """Cheap change detector: in-place updates (optimizer.step, load_state_dict) bump Tensor._version."""
            tr = self.__dict__.get('_tracked')
            if tr is None:
                tr = list(self.parameters()) + list(self.buffers())
                self.__dict__['_tracked'] = tr
            flat_ver = self._flat.version if self.__dict__.get('_flat') is not None else 0
            return sum(t._version for t in tr) + (1 << 40) * int(self._double_head) + (1 << 20) * flat_ver

        # ---- training ------------------------------------------------------------------------
        def flat_params(self):
            """Parameters re-homed into one flat fp32 buffer (+ flat gradient / bf16 shadow), see tdeed_b200/optim.py."""
            from tdeed_b200.optim import FlatParams
            if self._flat is None or not self._flat.valid():
                self._flat = FlatParams(self)
                self._train_engines.clear()
                self._train_graphs.clear()
            return self._flat

        def sync_replicas(self, force=False):
            """Data-parallel training (one process per GPU, torch.distributed initialised): broadcast rank 0's parameters and
            buffers once, so that replicas — whose temp_enc / SGP / heads / backbone were drawn from each rank's own RNG — apply
            the averaged gradients to identical weights.  Per-rank data / augmentation / mixup seeds remain the
            caller's job (tools/train_ddp.py offsets them by the rank).  BatchNorm running statistics stay per replica during
            training, as in the reference's single-GPU semantics per shard."""
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
                return
            if self.__dict__.get('_replicas_synced') and not force:
                return
            from tdeed_b200.parallel import broadcast_model
            broadcast_model(self)
            self._engines.clear()
            self.__dict__['_replicas_synced'] = True

        def train_engine(self, precision):
            from tdeed_b200.train_engine import TrainEngine
            flat = self.flat_params()
            eng = self._train_engines.get(precision)
            if eng is None:
                adt = torch.bfloat16 if precision == 'bf16' else torch.float32
                buffers = dict(self.named_buffers())
                eng = TrainEngine(self.engine_config(), flat.P, buffers, flat.G, act_dtype=adt,
                                  shadow=flat.S if precision == 'bf16' else None)
                self._train_engines[precision] = eng
            return eng

        def train_step(self, frame, label, labelD=None, fg_weight=5, grad_scale=1.0, accumulate=False, precision='bf16',
                       dropout_p=None, use_graph=None, dataset=None):
            """Forward (train mode) + loss + backward on the sm_100a kernels (model/model.py:262-324 of the reference).
            frame: (B,T,3,H,W) uint8 | float valued 0..255, already mixed up.  label: int64 (B*T) | float (B*T, K).
            Gradients of grad_scale * loss land in p.grad (added to the existing ones when accumulate).  Returns the
            loss as a device tensor [3] = (total, CE, MSE) — no host sync."""
            from tdeed_b200 import train_ops as TO
            flat = self.flat_params()
            self.sync_replicas()
            eng = self.train_engine(precision)
            if precision == 'bf16':
                flat.refresh_shadow()
            b, t, _, H, W = frame.shape
            # crop: same random window for the whole batch (RandomCrop on the flattened batch, model/model.py:115)
            if self.croping is not None:
                ch = cw = self.croping
                cy = int(torch.randint(0, H - ch + 1, (1,)).item()) if H > ch else 0
                cx = int(torch.randint(0, W - cw + 1, (1,)).item()) if W > cw else 0
            else:
                cy, cx, ch, cw = 0, 0, H, W
            unit = False
            if not isinstance(self.augmentation, nn.Identity):
                if self.fused_augmentation and frame.dtype in (torch.uint8, torch.float32):
                    # same transforms / distributions as self.augmentation, three fused kernel passes per clip (train_aug.cu)
                    from tdeed_b200.augment import ClipAugment
                    x = ClipAugment.apply(frame.contiguous(), (cy, cx, ch, cw), [ClipAugment.sample() for _ in range(b)])
                else:                  # the reference's torchvision pipeline, verbatim
                    x = frame[..., cy:cy + ch, cx:cx + cw].float() / 255.
                    for i in range(b):
                        x[i] = self.augmentation(x[i])
                frame, unit, cy, cx = x.contiguous(), True, 0, 0
            elif frame.dtype not in (torch.uint8, torch.float32):
                frame = frame.float()
            if (cy or cx) and (use_graph if use_graph is not None else self.use_train_graph):
                # a random crop window would make every step a new graph signature: cut the window out first
                frame, cy, cx = frame[..., cy:cy + ch, cx:cx + cw], 0, 0
            frame = frame.contiguous()
            self._train_calls += 1
            if dropout_p is None:            # nn.Dropout() of the FC heads (model/modules.py:366-376): p = 0.5 in train mode
                head = self._pred_fine._fc1 if self._double_head else self._pred_fine
                dropout_p = float(head.dropout.p) if self.training else 0.0
            hard = label.reshape(-1).contiguous() if not label.dtype.is_floating_point else None
            soft = label.float().contiguous() if label.dtype.is_floating_point else None
            use_d = labelD is not None and self._radi_displacement > 0
            labD = labelD.reshape(-1).float().contiguous() if use_d else None

            ds_dev = None
            if self._double_head:
                if dataset is None:
                    raise ValueError("double-head training needs batch['dataset'] (1 | 2 per clip)")
                ds_dev = torch.as_tensor(dataset, dtype=torch.int32).reshape(-1).to(frame.device)

            def run_a(fr, hd, sf, ld, grads):
                """forward + loss + backward of the heads and the temporal stack"""
                eng.G = grads
                logits, displ = eng.forward(fr, (cy, cx, ch, cw), unit_input=unit, dropout_p=dropout_p)
                loss = eng.loss(logits, displ, hd, sf, ld, fg_weight=fg_weight, dataset=ds_dev)
                eng.backward_temporal()
                return loss, logits, displ

            def run(fr, hd, sf, ld, grads):
                out = run_a(fr, hd, sf, ld, grads)
                eng.backward_backbone()
                return out

            direct = not accumulate and grad_scale == 1.0
            # data parallel (one process per GPU): the gradients of the temporal stack + heads — the tail of the flat buffer, ~80 %
            # of its bytes — are all-reduced over NVLink while the backbone backward still runs; the rest follows at the end
            import torch.distributed as dist
            overlap = (direct and self.overlap_allreduce and dist.is_available() and dist.is_initialized()
                       and dist.get_world_size() > 1)
            if overlap:
                from tdeed_b200.parallel import GradReducer, temporal_grad_range
                reducer = self.__dict__.setdefault('_reducer', GradReducer())
                t_lo, t_hi = temporal_grad_range(flat)
            if use_graph is None:
                use_graph = self.use_train_graph and not self._double_head     # (dataset ids are not a static graph input yet)
            if use_graph:
                # CUDA graph of forward + loss + backward for this (shapes, crop, dtypes) signature: ~3500 launches replayed
                # with one host call.  Inputs are copied into static buffers; gradients land in a static buffer.
                key = (precision, tuple(frame.shape), frame.dtype, (cy, cx, ch, cw), unit, hard is not None, use_d, dropout_p,
                       fg_weight, flat.p.data_ptr(), overlap)
                ent = self._train_graphs.get(key)
                if ent is None:                      # first sight of this signature: run eagerly (warms every kernel / allocator)
                    self._train_graphs[key] = 'warm'
                    use_graph = False
                elif ent == 'warm':
                    st = dict(frame=frame.clone(), hard=hard.clone() if hard is not None else None,
                              soft=soft.clone() if soft is not None else None, labD=labD.clone() if labD is not None else None,
                              g=torch.zeros_like(flat.g))
                    st['G'] = {n: st['g'][o:o + cnt].view(flat.P[n].shape) for n, (o, cnt) in flat.offsets.items()}
                    torch.cuda.synchronize()
                    # thread_local capture: the DataLoader's pin-memory thread and the prefetch stream may touch CUDA meanwhile
                    # (ADVICE r1); a failed capture falls back to eager execution for this signature
                    try:
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                            st['out'] = (run_a if overlap else run)(st['frame'], st['hard'], st['soft'], st['labD'], st['G'])
                        st['graph'] = graph
                        if overlap:         # second graph (same memory pool): the backbone backward, replayed after the first all-reduce is launched
                            graph_b = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(graph_b, pool=graph.pool(), capture_error_mode='thread_local'):
                                eng.backward_backbone()
                            st['graph_b'] = graph_b
                        self._train_graphs[key] = ent = st
                    except RuntimeError as exc:
                        import warnings
                        warnings.warn('tdeed_b200: CUDA-graph capture of the training step failed (%s); running eagerly' % (exc,))
                        self._train_graphs[key] = ent = 'eager'
                        eng.tape = None
                        use_graph = False
                elif ent == 'eager':
                    use_graph = False
                if use_graph:
                    ent['frame'].copy_(frame)
                    if hard is not None:
                        ent['hard'].copy_(hard)
                    if soft is not None:
                        ent['soft'].copy_(soft)
                    if labD is not None:
                        ent['labD'].copy_(labD)
                    ent['graph'].replay()
                    if 'graph_b' in ent:
                        reducer.launch(ent['g'], t_lo, t_hi)
                        ent['graph_b'].replay()
                        reducer.launch(ent['g'], 0, t_lo)
                        reducer.wait()
                        self._grads_reduced = True
                    loss, logits, displ = ent['out']
                    if accumulate:
                        TO.axpy_(ent['g'], grad_scale, flat.g)
                    elif grad_scale == 1.0:
                        flat.g.copy_(ent['g'])
                    else:
                        flat.g.zero_()
                        TO.axpy_(ent['g'], grad_scale, flat.g)
            if not use_graph:
                if overlap:
                    loss, logits, displ = run_a(frame, hard, soft, labD, flat.G)
                    reducer.launch(flat.g, t_lo, t_hi)
                    eng.backward_backbone()
                    reducer.launch(flat.g, 0, t_lo)
                    reducer.wait()
                    self._grads_reduced = True
                elif direct:
                    loss, logits, displ = run(frame, hard, soft, labD, flat.G)
                else:                               # accumulate / scale: write into a scratch buffer, then axpy
                    scratch = torch.zeros_like(flat.g)
                    G = {n: scratch[o:o + cnt].view(flat.P[n].shape) for n, (o, cnt) in flat.offsets.items()}
                    loss, logits, displ = run(frame, hard, soft, labD, G)
                    if not accumulate:
                        flat.g.zero_()
                    TO.axpy_(scratch, grad_scale, flat.g)
            flat.attach_grads()
            self._last_train = (logits.view(b, t, -1), displ.view(b, t) if displ is not None else None)
            return loss

        def engine(self, precision):
            """InferenceEngine for the current weights ('bf16' | 'fp32'); re-prepared when parameters changed."""
            dev = self.temp_enc.device
            if dev.type != 'cuda':
                raise RuntimeError('tdeed_b200 has no CPU path: move the model to a CUDA device (got %s)' % dev)
            ver = self._weights_version()
            eng = self._engines.get(precision)
            if eng is None:
                eng = InferenceEngine(self.engine_config(), self.state_dict(), device=dev, precision=precision)
                self._engines[precision] = eng
            elif self._engine_versions.get(precision) != ver:
                eng.cfg = self.engine_config()
                eng.load_state(self.state_dict())
            self._engine_versions[precision] = ver
            return eng

        # ---- forward -------------------------------------------------------------------------
        def forward(self, x, y=None, inference=False, augment_inference=False, use_graph=False):
            """x: (B, T, 3, H, W) valued 0..255 (uint8 or float).  Returns what the reference returns:
            ({'im_feat': logits, 'displ_feat': displ}, y) when radi_displacement > 0 else (logits, y)."""
            if not inference:
                raise RuntimeError(
                    'tdeed_b200: the train-mode forward has no torch.autograd graph — the backward pass is hand-written '
                    'sm_100a kernels.  Use Impl.train_step(frame, label, labelD) (what TDEEDModel.epoch(optimizer=...) '
                    'calls): it runs forward + loss + backward and leaves the gradients in p.grad')
            precision = 'bf16' if torch.is_autocast_enabled() else 'fp32'
            eng = self.engine(precision)
            if x.dtype not in (torch.uint8, torch.float32):
                x = x.float()
            x = x.contiguous()
            fwd = eng.forward_graphed if use_graph else eng.forward
            logits, displ, probs = fwd(x, flip=augment_inference)
            self._last_probs = probs
            if self._radi_displacement > 0:
                return {'im_feat': logits, 'displ_feat': displ}, y
            return logits, y

        def update_pred_head(self, num_classes=[1, 1]):
            self._pred_fine = FC2Layers(self._feat_dim, num_classes).to(self.temp_enc.device)
            self._double_head = True
            self._engines.clear()
            self.__dict__.pop('_tracked', None)
            self._flat = None                  # the parameter set changed: re-home on the next training step
            self._train_engines.clear()
            self._train_graphs.clear()

        def print_stats(self):
            print('Model params:', sum(p.numel() for p in self.parameters()))
            print('  CNN features:', sum(p.numel() for p in self._features.parameters()))
            print('  Temporal:', sum(p.numel() for p in self._temp_fine.parameters()))
            print('  Head:', sum(p.numel() for p in self._pred_fine.parameters()))

    train_precision = 'bf16'     # 'fp32' runs the exact CUDA-core kernels (parity mode)

    def _sync_gradients(self, optimizer):
        """Data-parallel training (one process per GPU): NCCL all-reduce of the flat gradient buffer over NVLink; the mean
        is folded into the fused AdamW kernel (grad_scale = 1 / world).  No-op for a single process."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        from tdeed_b200.parallel import allreduce_gradients
        flat = self._model.flat_params()
        if self._model._grads_reduced:          # already summed, overlapped with the backward pass (Impl.train_step)
            self._model._grads_reduced = False
            scale = 1.0 / dist.get_world_size()
        else:
            scale = allreduce_gradients(flat.g)
        if hasattr(optimizer, 'grad_scale'):
            optimizer.grad_scale = scale
        else:
            flat.g.mul_(scale)

    def __init__(self, device='cuda', args=None):
        self.device = device
        self._model = TDEEDModel.Impl(args=args)
        self._model.print_stats()
        self._args = args
        self._model.to(device)
        self._num_classes = args.num_classes + 1

    def epoch(self, loader, optimizer=None, scaler=None, lr_scheduler=None, acc_grad_iter=1, fg_weight=5,
              valMAP=False):
        """Same contract as the reference (model/model.py:193-332).  Evaluation (optimizer=None) runs on the
        sm_100a inference engine; the loss glue (cross-entropy / MSE on (B*T, K) logits) is host-side torch."""
        if optimizer is None:
            inference = True
            self._model.eval()
        else:
            inference = False
            optimizer.zero_grad()
            self._model.train()

        if valMAP:
            map_labels = []
            map_preds = []

        ce_kwargs = {}
        if fg_weight != 1:
            ce_kwargs['weight'] = torch.FloatTensor([1] + [fg_weight] * (self._num_classes - 1)).to(self.device)

        epoch_loss = 0.
        epoch_loss_dev = None
        with torch.no_grad() if optimizer is None else nullcontext():
            for batch_idx, batch in enumerate(tqdm(_DevicePrefetch(loader, self.device) if optimizer is not None else loader)):
                frame = batch['frame'].to(self.device, non_blocking=True)          # kept uint8: the stem kernel normalises
                label = batch['label'].to(self.device, non_blocking=True)

                if self._model._double_head:
                    batch_dataset = batch['dataset']
                    label = update_labels_2heads(label, batch_dataset, self._args.num_classes)

                if 'labelD' in batch.keys():
                    labelD = batch['labelD'].to(self.device, non_blocking=True).float()

                if 'frame2' in batch.keys():
                    # mixup (model/model.py:228-254 of the reference), same arithmetic — fl32(l)*frame + fl32(1-l)*frame2 and the
                    # two label adds in the same order — but batched and free of host syncs (the reference's per-clip
                    # `label_dist[i, range(T), label[i]] += l[i]` uploads an index tensor from pageable memory every clip,
                    # which drains the stream)
                    frame2 = batch['frame2'].to(self.device, non_blocking=True)
                    label2 = batch['label2'].to(self.device, non_blocking=True)
                    nb = frame2.shape[0]
                    l = [random.betavariate(0.2, 0.2) for _ in range(nb)]
                    lam = torch.tensor([[v, 1 - v] for v in l], dtype=torch.float64).float().pin_memory().to(self.device, non_blocking=True)
                    la, lb = lam[:, 0], lam[:, 1]
                    if (frame.dtype == torch.uint8 and frame2.dtype == torch.uint8 and frame.is_contiguous() and frame2.is_contiguous()
                            and frame[0].numel() % 16 == 0 and os.environ.get('TDEED_TORCH_MIXUP') != '1'):
                        from tdeed_b200 import train_ops
                        frame = train_ops.mixup_u8(frame, frame2, lam)        # one pass over the two uint8 clips (tdeed_mixup_u8)
                    else:
                        frame = la.view(nb, 1, 1, 1, 1) * frame.float() + lb.view(nb, 1, 1, 1, 1) * frame2.float()
                    label_dist = torch.zeros((label.shape[0], label.shape[1], self._num_classes), device=self.device)
                    label_dist.scatter_add_(2, label.unsqueeze(2), la.view(nb, 1, 1).expand(-1, label.shape[1], 1).contiguous())
                    label_dist.scatter_add_(2, label2.unsqueeze(2), lb.view(nb, 1, 1).expand(-1, label.shape[1], 1).contiguous())
                    if 'labelD2' in batch.keys():
                        labelD2 = batch['labelD2'].to(self.device, non_blocking=True).float()
                        labelD = la.view(nb, 1) * labelD + lb.view(nb, 1) * labelD2
                    label = label_dist

                if valMAP:
                    labels_aux = process_labels(label, labelD if 'labelD' in batch.keys() else None,
                                                num_classes=self._num_classes)
                    map_labels.append(labels_aux.cpu())

                label = label.flatten() if len(label.shape) == 2 else label.view(-1, label.shape[-1])

                if optimizer is not None:
                    # training: forward (train mode) + loss + backward are sm_100a kernels (tdeed_b200/train_engine.py);
                    # the reference's `step(optimizer, scaler, loss / acc_grad_iter, ...)` (model/modules.py:388-401)
                    # becomes gradient accumulation into p.grad + optimizer.step()
                    first = batch_idx % acc_grad_iter == 0
                    loss_dev = self._model.train_step(frame, label, labelD if 'labelD' in batch.keys() else None,
                                                      fg_weight=fg_weight, grad_scale=1.0 / acc_grad_iter,
                                                      accumulate=not first, precision=self.train_precision,
                                                      dataset=batch_dataset if self._model._double_head else None)
                    if valMAP:
                        logits, displ = self._model._last_train
                        map_preds.append((process_prediction(logits, displ) if displ is not None
                                          else torch.softmax(logits, dim=2)).cpu())
                    if (batch_idx + 1) % acc_grad_iter == 0:
                        self._sync_gradients(optimizer)
                        if scaler is None:
                            optimizer.step()
                        else:
                            scaler.step(optimizer)
                            scaler.update()
                        if lr_scheduler is not None:
                            lr_scheduler.step()
                        optimizer.zero_grad()
                    # no host sync per step: the loss stays on the device until the end of the epoch
                    epoch_loss_dev = loss_dev[0].clone() if epoch_loss_dev is None else epoch_loss_dev + loss_dev[0]
                    continue

                with torch.autocast('cuda', dtype=torch.bfloat16):
                    pred, y = self._model(frame, y=label, inference=inference)

                if 'labelD' in batch.keys():
                    predD = pred['displ_feat']
                    pred = pred['im_feat']

                if valMAP:
                    pred_aux = process_prediction(pred, predD)
                    map_preds.append(pred_aux.cpu())

                loss = 0.
                if self._model._double_head:
                    b, t, c = pred.shape
                    if len(label.shape) == 2:
                        label = label.view(b, t, c)
                    if len(label.shape) == 1:
                        label = label.view(b, t)
                    n1 = self._args.num_classes + 1
                    for i in range(pred.shape[0]):
                        if batch_dataset[i] == 1:
                            aux_label = label[i][:, :n1] if len(label.shape) == 3 else label[i]
                            loss += F.cross_entropy(pred[i][:, :n1], aux_label, weight=ce_kwargs['weight'][:n1]) / pred.shape[0]
                        elif batch_dataset[i] == 2:
                            aux_label = label[i][:, n1:] if len(label.shape) == 3 else label[i] - n1
                            loss += F.cross_entropy(pred[i][:, n1:], aux_label,
                                                    weight=ce_kwargs['weight'][:self._args.pretrain['num_classes'] + 1]) / pred.shape[0]
                else:
                    loss += F.cross_entropy(pred.reshape(-1, self._num_classes), label, **ce_kwargs)

                if 'labelD' in batch.keys():
                    loss = loss + F.mse_loss(predD, labelD, reduction='none').mean()

                if optimizer is not None:
                    step(optimizer, scaler, loss / acc_grad_iter, lr_scheduler=lr_scheduler,
                         backward_only=(batch_idx + 1) % acc_grad_iter != 0)

                epoch_loss += loss.detach().item()

        if epoch_loss_dev is not None:
            epoch_loss += float(epoch_loss_dev)
        if valMAP:
            return epoch_loss / len(loader), torch.cat(map_labels, 0), torch.cat(map_preds, 0)
        return epoch_loss / len(loader)

    def predict(self, seq, use_amp=True, augment_inference=False, use_graph=True):
        """(L,C,H,W) or (B,L,C,H,W) frames valued 0..255 -> (argmax (B,T) int64, probs (B,T,K) float32) ndarrays.
        use_amp=True runs the bf16 tensor-core engine (the reference: fp16 autocast), False the exact fp32 one.
        uint8 input is uploaded as uint8 (4x less H2D than the reference's float path) and normalised in the stem kernel."""
        if not isinstance(seq, torch.Tensor):
            seq = torch.as_tensor(np.asarray(seq))
        if len(seq.shape) == 4:
            seq = seq.unsqueeze(0)
        if seq.dtype not in (torch.uint8, torch.float32):
            seq = seq.float()
        if seq.device.type != 'cuda':
            seq = seq.to(self.device, non_blocking=True)

        self._model.eval()
        with torch.no_grad():
            with torch.autocast('cuda', dtype=torch.bfloat16) if use_amp else nullcontext():
                self._model(seq, inference=True, augment_inference=augment_inference, use_graph=use_graph)
            # softmax (+ displacement scatter-max over the first head) is fused into the heads kernel
            pred = self._model._last_probs.cpu().numpy()
        return np.argmax(pred, axis=2), pred


class _DevicePrefetch:
    """Iterates a loader one batch ahead and uploads the tensors of the NEXT batch on a side stream (pinned host memory ->
    non-blocking H2D), so that the copy overlaps the kernels of the current training step.  Non-tensor entries (e.g. the
    'dataset' list) pass through.  The reference blocks on `.to(device)` at the top of every iteration (model/model.py:216-234)."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, device

    def __len__(self):
        return len(self.loader)

    def _upload(self, batch, stream):
        if not torch.cuda.is_available() or torch.device(self.device).type != 'cuda':
            return batch, None
        out = {}
        with torch.cuda.stream(stream):
            for k, v in batch.items():
                out[k] = v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v
            ev = torch.cuda.Event()
            ev.record(stream)
        return out, ev

    def __iter__(self):
        stream = torch.cuda.Stream(device=self.device) if torch.cuda.is_available() else None
        it = iter(self.loader)
        try:
            nxt = self._upload(next(it), stream)
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._upload(next(it), stream)
            except StopIteration:
                nxt = None
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
                for v in cur.values():
                    if isinstance(v, torch.Tensor) and v.is_cuda:
                        v.record_stream(torch.cuda.current_stream())
            yield cur


def update_labels_2heads(labels, datasets, num_classes1=1):
    for i in range(len(datasets)):
        if datasets[i] == 2:
            labels[i] = labels[i] + num_classes1 + 1
    return labels
