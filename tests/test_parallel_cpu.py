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
