"""Multi-GPU inference plumbing: one process per GPU (torchrun), clips sharded by VIDEO, no data-path collective.

Clips are independent through the network (gate-shift and SGP never cross clip boundaries); the only cross-clip
coupling is the per-video averaging and NMS (util/eval.py:316-317,195-261 of the reference).  Giving each rank
whole videos keeps the fp32 accumulation order — and therefore every event — bit-identical to the single-GPU run
(SURVEY §8e).  The only exchange is a final all_gather of the (KB-sized) event lists.
"""
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_videos(videos, rank, world_size):
    """videos: iterable of (name, num_clips).  Deterministic longest-processing-time assignment of whole videos to
    ranks; returns the set of names owned by `rank`."""
    loads = [0] * world_size
    owner = {}
    for name, n in sorted(videos, key=lambda v: (-v[1], v[0])):
        r = min(range(world_size), key=lambda i: (loads[i], i))
        loads[r] += n
        owner[name] = r
    return {name for name, r in owner.items() if r == rank}


def gather_video_results(local, key=lambda d: d['video']):
    """local: list of per-video result dicts of this rank -> the merged list of ALL ranks, sorted by video name
    (the order `sorted(pred_dict.items())` gives in the reference), identical on every rank."""
    rank, ws = world()
    if ws == 1:
        return sorted(local, key=key)
    parts = [None] * ws
    dist.all_gather_object(parts, local)
    merged = [d for part in parts for d in part]
    return sorted(merged, key=key)
