"""World-size-2 gloo test of the multi-GPU plumbing (video sharding + result gather), run on CPU."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from tdeed_b200.parallel import gather_video_results, shard_videos


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


VIDEOS = [('v%02d' % i, n) for i, n in enumerate([169, 40, 12, 300, 7, 169, 88, 1, 25, 64, 64])]


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    mine = shard_videos(VIDEOS, rank, ws)
    local = [{'video': v, 'events': [{'frame': i, 'label': 'a', 'score': 0.5} for i in range(n % 5)], 'rank': rank}
             for v, n in VIDEOS if v in mine]
    merged = gather_video_results(local)
    out[rank] = ([d['video'] for d in merged], sorted(mine), sum(n for v, n in VIDEOS if v in mine))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(ws, _free_port(), out), nprocs=ws, join=True)
    names = sorted(v for v, _ in VIDEOS)
    assert out[0][0] == names and out[1][0] == names              # every rank sees every video, in name order
    assert sorted(out[0][1] + out[1][1]) == names                  # shards partition the videos
    assert not set(out[0][1]) & set(out[1][1])
    total = sum(n for _, n in VIDEOS)
    assert abs(out[0][2] - out[1][2]) <= max(n for _, n in VIDEOS) and out[0][2] + out[1][2] == total


def test_shard_is_deterministic_and_balanced():
    for ws in (1, 2, 4, 8):
        shards = [shard_videos(VIDEOS, r, ws) for r in range(ws)]
        assert shards == [shard_videos(list(reversed(VIDEOS)), r, ws) for r in range(ws)]
        assert sorted(v for s in shards for v in s) == sorted(v for v, _ in VIDEOS)
        loads = [sum(n for v, n in VIDEOS if v in s) for s in shards]
        assert max(loads) - min(loads) <= 300


def _grad_worker(rank, ws, port, out):
    import torch
    from tdeed_b200.parallel import allreduce_gradients, broadcast_parameters
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    scale = allreduce_gradients(g, bucket_elems=300)              # 4 buckets, the last one ragged
    p = torch.full((10,), float(rank + 5))
    broadcast_parameters(p, src=0)
    out[rank] = (g.tolist(), scale, p.tolist())
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_world2():
    """DP training plumbing: bucketed all-reduce of the flat gradient (sum) + the 1/world factor for the fused AdamW."""
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_grad_worker, args=(ws, _free_port(), out), nprocs=ws, join=True)
    want = [3.0 * i for i in range(1000)]
    for r in range(ws):
        g, scale, p = out[r]
        assert g == want and scale == 0.5 and p == [5.0] * 10


def test_bind_to_gpu_numa_is_harmless_without_a_gpu():
    """The per-rank CPU binding used by the multi-rank bench must never raise and must leave the affinity alone when NVML (or
    the device) is not there."""
    from tdeed_b200.parallel import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    cpus = bind_to_gpu_numa(0)
    after = os.sched_getaffinity(0)
    assert cpus is None or set(cpus) == after
    if cpus is None:
        assert after == before
    os.sched_setaffinity(0, before)


def _reducer_worker(rank, ws, port, out):
    import torch
    from tdeed_b200.parallel import GradReducer
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    red = GradReducer(bucket_elems=128)
    red.launch(g, 600, 1000)            # the tail (temporal stack + heads) first, while "the backbone backward still runs" ...
    g[:600] += 1.0                      # ... and keeps writing the head of the buffer
    red.launch(g, 0, 600)
    scale = red.wait()
    out[rank] = (g.tolist(), scale, red.handles, red.reduced)
    dist.destroy_process_group()


def test_grad_reducer_ranges_world2():
    """Overlapped data-parallel reduction: ranges launched at different times cover the flat buffer exactly once."""
    ws = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_reducer_worker, args=(ws, _free_port(), out), nprocs=ws, join=True)
    want = [3.0 * i + (2.0 if i < 600 else 0.0) for i in range(1000)]
    for r in range(ws):
        g, scale, handles, reduced = out[r]
        assert g == want and scale == 0.5 and handles == [] and reduced == []


def test_temporal_grad_range_is_the_tail_of_the_flat_buffer():
    """The parameters finished by TrainEngine.backward_temporal() form one contiguous tail range of FlatParams' layout."""
    from collections import OrderedDict
    from types import SimpleNamespace
    from tdeed_b200.parallel import temporal_grad_range
    import tdeed_oracle as O
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=8, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=3, radi_displacement=1, crop_dim=None)
    offs, total = OrderedDict(), 0
    for name, shape in O.state_shapes(cfg).items():          # == named_parameters() order minus the BN buffers
        if name.endswith(('running_mean', 'running_var', 'num_batches_tracked')):
            continue
        n = 1
        for v in shape:
            n *= v
        offs[name] = (total, n)
        total += (n + 15) // 16 * 16
    lo, hi = temporal_grad_range(SimpleNamespace(offsets=offs, total=total))
    assert hi == total and lo == offs['_temp_fine._sgp.0.ln.weight'][0]
    tail = sum(n for name, (o, n) in offs.items() if o >= lo)
    assert tail / sum(n for _, n in offs.values()) > 0.7      # SGP + heads dominate the parameter count
    assert all(name.startswith(('_temp_fine.', '_pred_')) for name, (o, _) in offs.items() if o >= lo)


def test_rank_loader_reseeds_workers_per_rank():
    """tools/train_ddp.py: train_tdeed.py:126-127 seeds worker i with `i + epoch * 100` on EVERY rank; the launcher's DataLoader
    subclass must give different ranks different sample streams and keep each rank reproducible."""
    import importlib.util
    import random
    import torch
    from torch.utils.data import DataLoader, Dataset
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('train_ddp', os.path.join(root, 'tools', 'train_ddp.py'))
    ddp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ddp)

    class Draw(Dataset):
        def __len__(self):
            return 8

        def __getitem__(self, i):
            return torch.tensor([random.random(), float(torch.rand(1))])

    def reference_init(worker_id):          # train_tdeed.py:126-127 with epoch = 0
        random.seed(worker_id)

    def stream(rank):
        cls = ddp.make_rank_loader(DataLoader, rank)
        return torch.cat([b for b in cls(Draw(), batch_size=4, num_workers=2, worker_init_fn=reference_init)]).tolist()
    a0, a0_again, a1 = stream(0), stream(0), stream(1)
    assert a0 == a0_again and a0 != a1
    plain = torch.cat([b for b in DataLoader(Draw(), batch_size=4, num_workers=2, worker_init_fn=reference_init)])[:, 0].tolist()
    assert plain != [v[0] for v in a0]
    assert len({ddp.rank_seed(1, r) for r in range(8)}) == 8
