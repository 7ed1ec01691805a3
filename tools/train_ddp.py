#!/usr/bin/env python3
"""Data-parallel launcher for the reference's UNMODIFIED train_tdeed.py (one process per GPU, NCCL over NVLink):

    cd <T-DEED checkout>
    PYTHONPATH=<this repo>/t-deed_b200:$PYTHONPATH \\
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        <this repo>/tools/train_ddp.py --model FineGym_big --seed 1

What the wrapper adds around `train_tdeed.main(get_args())` — nothing inside it changes:
  * process group + device + NUMA binding (tdeed_b200.parallel.init_distributed); TDEEDModel then broadcasts rank 0's weights
    and BN buffers before the first step (Impl.sync_replicas), all-reduces the flat gradient buffer — the temporal-stack part
    overlapped with the backbone backward — and averages inside the fused AdamW kernel (model/model.py of this repo);
  * per-rank sampling: train_tdeed.py:126-127 seeds every DataLoader worker with `id + epoch * 100`, which would give every
    rank the SAME clips, mixup partners and augmentations.  The DataLoader class train_tdeed imported is replaced by a
    subclass that re-seeds `random` / `numpy.random` / `torch` in each worker from (the reference's seed, the rank), and the
    process-level seeds of train_tdeed.py:93-95 are offset by the rank after main() sets them (mixup lambdas, crop windows);
  * single writer: checkpoints (`torch.save`), loss.json and wandb calls only happen on rank 0;
  * the global batch is `batch_size` x world (weak scaling, config batch per GPU); BatchNorm statistics stay per
    replica.  Evaluation at the end is sharded by video inside util.eval.evaluate and merged on every rank.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 't-deed_b200'))

from tdeed_b200.parallel import init_distributed  # noqa: E402

_MIX = 0x9E3779B97F4A7C15


def rank_seed(base, rank):
    """Deterministic per-rank seed derived from a base seed (64-bit mix, then folded to 32 bits for numpy)."""
    v = (int(base) * 6364136223846793005 + (rank + 1) * _MIX) & 0xFFFFFFFFFFFFFFFF
    return (v ^ (v >> 29)) & 0xFFFFFFFF


def make_rank_loader(base_cls, rank):
    """DataLoader subclass whose workers are re-seeded per rank AFTER the caller's own worker_init_fn ran."""

    class RankDataLoader(base_cls):
        def __init__(self, *args, worker_init_fn=None, **kwargs):
            def init(worker_id, _inner=worker_init_fn):
                if _inner is not None:
                    _inner(worker_id)
                base = random.getrandbits(48)          # deterministic: drawn right after the reference's own seeding
                s = rank_seed(base, rank)
                random.seed(s)
                np.random.seed(s)
                torch.manual_seed(s)
            super().__init__(*args, worker_init_fn=init, **kwargs)

    return RankDataLoader


def main():
    rank, world, local = init_distributed()
    if '' not in sys.path and os.getcwd() not in sys.path:
        sys.path.insert(1, os.getcwd())                 # the T-DEED checkout (train_tdeed.py, dataset/, util/io.py ...)
    import train_tdeed

    if world > 1:
        train_tdeed.DataLoader = make_rank_loader(train_tdeed.DataLoader, rank)
        set_seed = torch.manual_seed

        def manual_seed(seed):                          # train_tdeed.py:93: the first thing main() does
            out = set_seed(rank_seed(seed, rank))
            return out
        train_tdeed.torch.manual_seed = manual_seed
        np_seed, py_seed = np.random.seed, random.seed
        train_tdeed.np.random.seed = lambda s=None: np_seed(rank_seed(s, rank) if s is not None else None)
        train_tdeed.random.seed = lambda s=None, *a: py_seed(rank_seed(s, rank) if isinstance(s, int) else s)
        if rank != 0:
            train_tdeed.torch.save = lambda *a, **k: None
            train_tdeed.store_json = lambda *a, **k: None
            for name in ('login', 'init', 'log'):
                setattr(train_tdeed.wandb, name, lambda *a, **k: None)
            train_tdeed.wandb.summary = {}
    try:
        train_tdeed.main(train_tdeed.get_args())
    finally:
        if world > 1:
            torch.manual_seed, np.random.seed, random.seed = set_seed, np_seed, py_seed
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier()
                dist.destroy_process_group()


if __name__ == '__main__':
    main()
