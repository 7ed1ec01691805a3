"""Two-GPU checks (NCCL, one process per GPU; skipped on a single-GPU box):
  * data-parallel training: the all-reduced flat gradient (temporal-stack part overlapped with the backbone backward, eager and
    CUDA-graph replay) is EXACTLY g_rank0 + g_rank1 of the same deterministic kernels run single-process on each shard — i.e.
    reference-per-shard semantics with per-replica BatchNorm statistics and averaged gradients; replicas start from
    rank 0's weights;
  * clip-sharded `util.eval.evaluate`: every rank owns whole videos, the merged result (mAPs, printed tables, JSON files) is
    identical to the single-GPU run."""
import glob
import io
import json
import os
import socket
import tempfile
from argparse import Namespace
from contextlib import redirect_stdout

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')


def _setup(rank, ws, port):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, 't-deed_b200'), os.path.join(root, 'oracle')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    from tdeed_b200.parallel import init_distributed
    init_distributed()


def _model(cfg, seed, rank):
    import tdeed_oracle as O
    from model.model import TDEEDModel
    args = Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=cfg.radi_displacement,
                     feature_arch=cfg.feature_arch, clip_len=cfg.clip_len, n_layers=cfg.n_layers, sgp_ks=cfg.sgp_ks,
                     sgp_r=cfg.sgp_r, num_classes=cfg.num_classes, crop_dim=cfg.crop_dim)
    with redirect_stdout(io.StringIO()):
        m = TDEEDModel(device='cuda:%d' % rank, args=args)
    m.load(O.random_state(cfg, seed))
    return m


def _dp_worker(rank, ws, port, out):
    import torch.distributed as dist
    _setup(rank, ws, port)
    import tdeed_oracle as O
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=6, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=3, radi_displacement=1, crop_dim=None)
    m = _model(cfg, 7 + rank, rank)                    # DIFFERENT weights per rank on purpose: sync_replicas must fix that
    m._model.augmentation = torch.nn.Identity()
    m._model.train()
    g = torch.Generator().manual_seed(100 + rank)     # a different shard per rank
    frames = torch.randint(0, 256, (2, 6, 3, 32, 48), generator=g, dtype=torch.uint8).float().cuda()
    label = torch.randint(0, 4, (2, 6), generator=g).cuda().reshape(-1)
    labelD = torch.randint(-1, 2, (2, 6), generator=g).float().cuda()
    opt, _ = m.get_optimizer({'lr': 1e-3})             # builds the flat buffers and broadcasts rank 0's weights
    flat = m._model.flat_params()
    w = [torch.empty_like(flat.p) for _ in range(ws)]
    dist.all_gather(w, flat.p)
    same_weights = bool(torch.equal(w[0], w[1]))
    res = {'same_weights': same_weights}
    for precision in ('fp32', 'bf16'):
        # single-process gradient of this rank's shard (no reduction)
        m._model.overlap_allreduce = False
        m._model.train_step(frames, label, labelD, precision=precision, dropout_p=0.0, use_graph=False)
        mine = flat.g.clone()
        both = [torch.empty_like(mine) for _ in range(ws)]
        dist.all_gather(both, mine)
        want = both[0] + both[1]
        m._model.overlap_allreduce = True
        ok = []
        for it in range(4):                            # eager (1st: warm), then CUDA-graph replay from the 3rd call on
            m._model.train_step(frames, label, labelD, precision=precision, dropout_p=0.0)
            assert m._model._grads_reduced
            m._sync_gradients(opt)
            ok.append(bool(torch.equal(flat.g, want)))
        graphs = [v for v in m._model._train_graphs.values() if isinstance(v, dict)]
        res[precision] = dict(ok=ok, scale=opt.grad_scale, two_graphs=any('graph_b' in v for v in graphs),
                              nonzero=float(want.abs().sum()) > 0)
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradients_equal_sum_of_per_shard_gradients():
    _need_two_gpus()
    out = mp.Manager().dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        assert out[r]['same_weights']
        for precision in ('fp32', 'bf16'):
            d = out[r][precision]
            assert d['ok'] == [True] * 4 and d['scale'] == 0.5 and d['two_graphs'] and d['nonzero'], (r, precision, d)


def _eval_worker(rank, ws, port, out):
    import torch.distributed as dist
    _setup(rank, ws, port)
    import synth_data as S
    import tdeed_oracle as O
    import util.eval as E
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=16, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=4, radi_displacement=1, crop_dim=32)
    m = _model(cfg, 5, rank)
    classes = {'jump': 1, 'spin': 2, 'step': 3, 'fall': 4}
    lengths = {'a': 70, 'b': 45, 'c': 12, 'd': 90, 'e': 33}
    ds = S.SyntheticVideoDataset(classes, lengths=lengths, hw=(32, 56), clip_len=16, overlap_len=12, stride=1, dataset='fs_comp',
                                 seed=21, events_per_100=6.0)

    def run(tmp):
        buf = io.StringIO()
        save = os.path.join(tmp, 'run', 'pred-test')
        with redirect_stdout(buf):
            ret = E.evaluate(m, ds, 'TEST', classes, save, printed=True, test=True, augment=True)
        files = {os.path.relpath(p, tmp): open(p).read() for p in sorted(glob.glob(os.path.join(tmp, '**', '*.json'), recursive=True))}
        return [float(x) for x in ret[0]], files, buf.getvalue()
    with tempfile.TemporaryDirectory() as tmp:
        sharded = run(tmp) if rank == 0 else run(tempfile.mkdtemp())
    world = E.world
    E.world = lambda: (0, 1)                       # the same process, un-sharded
    try:
        with tempfile.TemporaryDirectory() as tmp:
            single = run(tmp)
    finally:
        E.world = world
    out[rank] = dict(maps_equal=sharded[0] == single[0], stdout_equal=sharded[2] == single[2],
                     files_equal=(sharded[1] == single[1]) if rank == 0 else True,
                     n_events=sum(len(v['events']) for v in json.loads(single[1]['run/pred-test.json'])))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluate_equals_single_gpu():
    _need_two_gpus()
    out = mp.Manager().dict()
    mp.spawn(_eval_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        assert out[r]['maps_equal'] and out[r]['stdout_equal'] and out[r]['files_equal'] and out[r]['n_events'] > 0, (r, dict(out[r]))
