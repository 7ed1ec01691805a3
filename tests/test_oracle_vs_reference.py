"""Build-container-only: the oracle against the UNMODIFIED reference imported from /root/reference."""
import numpy as np
import pytest
import torch

import tdeed_oracle as O
import postproc_oracle as P

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope='module')
def ref():
    from ref_import import reference_modules
    cm = reference_modules()
    mods = cm.__enter__()
    yield mods
    cm.__exit__(None, None, None)


def test_timm_shim_param_counts(ref):
    import timm
    assert sum(p.numel() for p in timm.create_model('regnety_002').parameters()) == 3162996
    assert sum(p.numel() for p in timm.create_model('regnety_008').parameters()) == 6263168


@pytest.mark.parametrize('name,params', [('FineDiving_small', 12267634), ('FigureSkatingComp_small', 16363474)])
def test_state_layout_and_param_count(ref, name, params):
    from gen_golden import build_reference_model
    cfg = O.named_config(name)
    sd = O.random_state(cfg, 0)
    model = build_reference_model(ref, cfg, sd)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys())
    assert sum(p.numel() for p in model._model.parameters()) == params


def test_sgp_block_and_mixer(ref):
    torch.manual_seed(0)
    M = ref['model.modules']
    for c, t, ks, r in [(368, 25, 5, 4), (64, 13, 9, 2)]:
        blk, mix = M.SGPBlock(c, kernel_size=ks, k=r, init_conv_vars=0.1), M.SGPMixer(c, kernel_size=ks, k=r, init_conv_vars=0.1, t_size=2 * t - 1)
        for m in (blk, mix):
            for p in m.parameters():
                p.data.add_(torch.randn_like(p) * 0.05)
        x, z = torch.randn(2, c, t), torch.randn(2, c, 2 * t - 1)
        sd = {'b.' + k: v for k, v in blk.state_dict().items()}
        sd.update({'m.' + k: v for k, v in mix.state_dict().items()})
        with torch.no_grad():
            assert torch.allclose(O.sgp_block(x, sd, 'b'), blk(x), atol=2e-5, rtol=1e-5)
            assert torch.allclose(O.sgp_mixer(x, z, sd, 'm', 2 * t - 1), mix(x, z), atol=2e-5, rtol=1e-5)


def test_scatter_max_vs_python_loop(ref):
    torch.manual_seed(1)
    M = ref['model.modules']
    logits, displ = torch.randn(3, 40, 7) * 2, torch.randn(3, 40) * 3
    displ[0, :6] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 60.0])     # half-even + clamp cases
    assert torch.equal(O.scatter_max_probs(logits, displ), M.process_prediction(logits, displ))
    assert torch.equal(O.scatter_max_probs(logits, displ, num_softmax=4), M.process_double_head(logits, displ, num_classes=4))


def test_nms_random_vs_reference(ref):
    ev = ref['util.eval']
    rng = np.random.default_rng(3)
    classes = {'a': 1, 'b': 2, 'c': 3}
    inv = {v: k for k, v in classes.items()}
    for trial in range(40):
        n = int(rng.integers(1, 120))
        fr = np.sort(rng.integers(0, 80, n)).astype(np.int32)
        lb = rng.integers(1, 4, n).astype(np.int32)
        keep = np.unique(np.stack([fr, lb], 1), axis=0, return_index=True)[1]
        fr, lb = fr[np.sort(keep)], lb[np.sort(keep)]
        sc = rng.choice(np.asarray([0.01, 0.2, 0.2, 0.5, 0.05, 0.009], np.float32), len(fr)) * \
            rng.choice(np.asarray([1, 1, 1, 0.999], np.float32), len(fr))
        sc = sc.astype(np.float32)
        w = int(rng.integers(1, 7))
        thr = float(rng.choice([0.0, 0.01, 0.1]))
        dicts = [P.to_dicts('v', fr, lb, sc, inv)]
        for mine, theirs in ((P.nms(fr, lb, sc, w, thr), ev.non_maximum_supression(dicts, w, thr)),
                             (P.soft_nms(fr, lb, sc, w, 0.01), ev.soft_non_maximum_supression(dicts, w, 0.01))):
            f, l, s = P.from_dicts(theirs[0], classes)
            assert np.array_equal(mine[0], f) and np.array_equal(mine[1], l)
            assert np.array_equal(mine[2].astype(np.float64), s)
