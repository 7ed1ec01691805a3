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


# ---------------------------------------------------------------------------------------------------------------------
# data-parallel training: replicas, per-replica BatchNorm statistics, gradients averaged over ranks
# ---------------------------------------------------------------------------------------------------------------------
def allreduce_gradients(flat_grad, bucket_elems=16 << 20):
    """Sum the flat gradient buffer over all ranks (NCCL over NVLink on the GPU box; gloo in the CPU tests), in buckets of
    `bucket_elems` elements issued asynchronously and waited on together.  Returns the factor that turns the sum into the
    mean (1 / world) — the fused AdamW kernel applies it (grad_scale), so no extra pass over the gradients is needed."""
    rank, ws = world()
    if ws == 1:
        return 1.0
    handles = []
    n = flat_grad.numel()
    for lo in range(0, n, bucket_elems):
        handles.append(dist.all_reduce(flat_grad[lo:min(n, lo + bucket_elems)], op=dist.ReduceOp.SUM, async_op=True))
    for h in handles:
        h.wait()
    return 1.0 / ws


class GradReducer:
    """All-reduce of RANGES of the flat gradient buffer, launched as soon as a range is final while the rest of the
    backward pass still runs: `launch()` enqueues bucketed async all-reduces (NCCL runs them on its own stream, ordered
    after the kernels already enqueued on the current stream), `wait()` makes the current stream wait for all of them and
    returns the sum -> mean factor 1 / world.  With one process both are no-ops."""

    def __init__(self, bucket_elems=16 << 20):
        self.bucket_elems = bucket_elems
        self.handles = []
        self.reduced = []              # [(lo, hi)] ranges launched since the last wait()

    def launch(self, flat_grad, lo, hi):
        rank, ws = world()
        self.reduced.append((lo, hi))
        if ws == 1:
            return
        for a in range(lo, hi, self.bucket_elems):
            self.handles.append(dist.all_reduce(flat_grad[a:min(hi, a + self.bucket_elems)], op=dist.ReduceOp.SUM, async_op=True))

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles = []
        self.reduced = []
        return 1.0 / world()[1]


def temporal_grad_range(flat):
    """[lo, hi) of the flat buffers holding the parameters whose gradients are final after TrainEngine.backward_temporal():
    everything from the first `_temp_fine.*` parameter to the end (`_temp_fine`, `_pred_fine`, `_pred_displ` are registered
    after the backbone in TDEEDModel.Impl.__init__, model/model.py:65-75 of the reference)."""
    names = list(flat.offsets)
    first = next(i for i, n in enumerate(names) if n.startswith('_temp_fine.'))
    assert all(not n.startswith(('_features.', 'temp_enc')) for n in names[first:]), 'unexpected parameter order'
    return flat.offsets[names[first]][0], flat.total


def broadcast_model(impl, src=0):
    """Make every replica identical to rank `src`: the flat parameter buffer (or every parameter when it has not been built
    yet) and all buffers (BatchNorm running statistics).  ADVICE r1: replicas initialise temp_enc / SGP / heads from their own
    RNG, and averaged gradients applied to different weights diverge."""
    rank, ws = world()
    if ws == 1:
        return
    flat = impl.__dict__.get('_flat')
    if flat is not None and flat.valid():
        dist.broadcast(flat.p, src=src)
        flat.version += 1
    else:
        for p in impl.parameters():
            dist.broadcast(p.data, src=src)
    for b in impl.buffers():
        dist.broadcast(b, src=src)


def init_distributed(backend=None):
    """One process per GPU under torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment): binds the process
    to its GPU (and that GPU's NUMA node), creates the process group.  Returns (rank, world, local_rank); (0, 1, 0) and no
    process group when WORLD_SIZE is unset or 1."""
    import os
    import torch
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if ws > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            bind_to_gpu_numa(local)
            dist.init_process_group(backend or 'nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend or 'gloo')
    return rank, ws, local


def broadcast_parameters(flat_param, src=0):
    """Make every replica start from rank `src`'s weights (flat parameter buffer, in place)."""
    rank, ws = world()
    if ws > 1:
        dist.broadcast(flat_param, src=src)


def bind_to_gpu_numa(device_index):
    """Restrict this process to the CPUs NVML reports as local to GPU `device_index` (its NUMA node / PCIe root complex), so that
    the pinned staging buffers it allocates afterwards (first touch) and the threads that fill them sit next to the GPU.  With
    eight ranks each streaming ~30 GB/s of pinned clips, un-pinned ranks share one socket's memory controllers and the
    end-to-end rate stops scaling (5.5x at 8 GPUs, r1g).  Returns the CPU list, or None when NVML / the affinity call is
    unavailable (the process is then left untouched).  Call it once per rank, before allocating pinned memory."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(device_index)
            bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
