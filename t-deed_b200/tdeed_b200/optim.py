"""Flat parameter storage + fused AdamW for the training step.

`FlatParams` re-homes every parameter of a module into ONE fp32 buffer (16-element aligned slices, so every slice is a
valid TMA / vector operand), with a same-shaped flat gradient buffer and a bf16 shadow for the tensor-core GEMM operands.
`FusedAdamW` is a real `torch.optim.Optimizer` (LR schedulers, state_dict, param_groups work as usual; the reference
builds `torch.optim.AdamW(self._get_params(), **opt_args)`, model/modules.py:37-39) whose step is one launch of
`tdeed_adamw_step` over the flat buffer instead of ~10 elementwise kernels per parameter tensor.
Data-parallel training all-reduces `FlatParams.g` (one NCCL call per bucket) — see tdeed_b200.parallel.
"""
from collections import OrderedDict

import torch

from . import train_ops as T

ALIGN = 16


class FlatParams:
    def __init__(self, module):
        named = [(n, p) for n, p in module.named_parameters()]
        dev = named[0][1].device
        if dev.type != 'cuda':
            raise RuntimeError('tdeed_b200 has no CPU path: move the model to a CUDA device before training')
        offs, total = [], 0
        for _, p in named:
            offs.append(total)
            total += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.names, self.params, self.offsets = [], [], OrderedDict()
        self.P, self.G, self.S = {}, {}, {}
        with torch.no_grad():
            for (n, p), o in zip(named, offs):
                sl = slice(o, o + p.numel())
                self.p[sl].copy_(p.data.reshape(-1))
                p.data = self.p[sl].view(p.shape)
                self.names.append(n)
                self.params.append(p)
                self.offsets[n] = (o, p.numel())
                self.P[n] = p.data
                self.G[n] = self.g[sl].view(p.shape)
                self.S[n] = self.shadow[sl].view(p.shape)
        self.total = total
        self.version = 0

    def valid(self):
        """False when something (module.to(), load with assign=True, ...) moved a parameter out of the flat buffer."""
        base = self.p.data_ptr()
        return all(p.data_ptr() == base + 4 * self.offsets[n][0] for n, p in zip(self.names, self.params))

    def refresh_shadow(self):
        T.cast(self.p, torch.bfloat16, out=self.shadow)

    def attach_grads(self):
        for n, p in zip(self.names, self.params):
            p.grad = self.G[n]


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, flat=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.flat = flat
        self.grad_scale = 1.0            # data-parallel: 1 / world_size (the all-reduce sums)

    def _flat_group(self, group):
        f = self.flat
        return (f is not None and len(self.param_groups) == 1 and len(group['params']) == len(f.params)
                and all(a is b for a, b in zip(group['params'], f.params)) and f.valid())

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            hp = (group['lr'], group['betas'], group['eps'], group['weight_decay'])
            if self._flat_group(group):
                if all(p.grad is None for p in group['params']):
                    continue
                st = self.state['flat']
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(self.flat.p)
                    st['exp_avg_sq'] = torch.zeros_like(self.flat.p)
                st['step'] += 1
                T.adamw_step_(self.flat.p, self.flat.g, st['exp_avg'], st['exp_avg_sq'], *hp, st['step'],
                              grad_scale=self.grad_scale, shadow=self.flat.shadow)
                self.flat.version += 1        # kernels write behind torch's back: Tensor._version does not move
                continue
            for p in group['params']:                     # generic path: one launch per tensor
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError('FusedAdamW: tdeed_b200 has no CPU path')
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p.data, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p.data, memory_format=torch.contiguous_format)
                st['step'] += 1
                T.adamw_step_(p.data, p.grad.contiguous(), st['exp_avg'], st['exp_avg_sq'], *hp, st['step'], grad_scale=self.grad_scale)
                p.add_(0)                                 # bump Tensor._version (the kernel wrote behind torch's back)
        return loss
