"""CPU-only checks of the host side: the C-ABI library exports what include/tdeed_b200.h declares, the
drop-in model package mirrors the reference's state_dict layout, and the product fails loudly without CUDA."""
import ctypes
import os
import re
from argparse import Namespace

import pytest
import torch

import tdeed_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(cfg):
    return Namespace(modality='rgb', temporal_arch='ed_sgp_mixer', radi_displacement=cfg.radi_displacement,
                     feature_arch=cfg.feature_arch, clip_len=cfg.clip_len, n_layers=cfg.n_layers, sgp_ks=cfg.sgp_ks,
                     sgp_r=cfg.sgp_r, num_classes=cfg.num_classes, crop_dim=cfg.crop_dim)


def test_library_exports_every_declared_symbol():
    from tdeed_b200 import _lib
    header = ''.join(open(os.path.join(ROOT, 'include', h)).read() for h in ('tdeed_b200.h', 'tdeed_b200_train.h'))
    declared = set(re.findall(r'\b(tdeed_[a-z0-9_]+)\s*\(', header))
    declared -= {'tdeed_gemm_seg', 'tdeed_status'}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    # argument counts of the ctypes prototypes == the header's
    for name, args in re.findall(r'\b(tdeed_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', re.sub(r'/\*.*?\*/', '', header, flags=re.S)):
        n = 0 if args.strip() in ('', 'void') else len(args.split(','))
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))
    lib = ctypes.CDLL(_lib.LIB_PATH)           # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().tdeed_abi_version() == 1


@pytest.mark.parametrize('name', ['FineDiving_small', 'FineGym_big', 'SoccerNetBall_challenge2'])
def test_dropin_state_dict_matches_reference_layout(name):
    from model.model import TDEEDModel
    cfg = O.named_config(name)
    m = TDEEDModel(device='cpu', args=_args(cfg))
    if cfg.double_head:
        m._model.update_pred_head(cfg.double_head)
    sd = m.state_dict()
    ref = O.state_shapes(cfg)
    assert list(sd.keys()) == list(ref.keys())
    for k, shape in ref.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    m.load(O.random_state(cfg, 0))             # strict load of a reference-layout checkpoint


def test_no_cpu_fallback():
    from model.model import TDEEDModel
    cfg = O.Config(clip_len=4, crop_dim=32)
    m = TDEEDModel(device='cpu', args=_args(cfg))
    with pytest.raises(RuntimeError, match='no CPU path'):
        m.predict(torch.zeros(1, 4, 3, 32, 32, dtype=torch.uint8), use_amp=False)


def test_gemm_rejects_bad_arguments_without_gpu():
    from tdeed_b200 import _lib as L
    lib = L.load()
    seg = (L.GemmSeg * 1)()
    seg[0].a, seg[0].lda, seg[0].col0, seg[0].k = 16, 8, 0, 8
    rc = lib.tdeed_gemm_fwd(L.F32, 4, 12, 1, seg, 1, 0, 0, 16, None, None, 0, 0, 0, 16, 16, 0, 0, None)
    assert rc == -1 and b'multiples of 8' in lib.tdeed_last_error()


def test_pretrained_backbone_is_loaded_strictly_or_warned_about(tmp_path, monkeypatch):
    """create_model(pretrained=True) must not silently return random weights (model/model.py:38-46 of the
    reference starts from timm's ImageNet weights)."""
    import warnings
    import torch
    from model import regnet
    monkeypatch.delenv(regnet.PRETRAINED_ENV, raising=False)
    with pytest.warns(RuntimeWarning, match='RANDOMLY initialised'):
        regnet.create_model('regnety_002', pretrained=True)
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        src = regnet.create_model('regnety_002', pretrained=False)          # no warning without pretrained
    with torch.no_grad():
        for p in src.parameters():
            p.normal_()
    torch.save(src.state_dict(), tmp_path / 'regnety_002.pth')
    monkeypatch.setenv(regnet.PRETRAINED_ENV, str(tmp_path))
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        got = regnet.create_model('regnety_002', pretrained=True)
    for (k, a), (_, b) in zip(src.state_dict().items(), got.state_dict().items()):
        assert torch.equal(a, b), k
    bad = {k: v for k, v in src.state_dict().items() if k != 'stem.conv.weight'}
    torch.save(bad, tmp_path / 'regnety_008.pth')
    with pytest.raises(RuntimeError):                                        # strict: wrong / missing keys raise
        regnet.create_model('regnety_008', pretrained=True)


def test_stage_width_padding_keeps_the_network_function():
    """engine.pad_block_state zero-pads a bottleneck (RegNetY-200MF stage 3: 152 -> 160 channels in the bf16 engine): the original
    tensors sit unchanged in the leading block, the pad channels get zero weights / BN scale and shift (weight 0, bias 0, mean 0,
    var 1) / SE columns — checked here on the CPU with the oracle: a padded block maps the same input to the same output."""
    from tdeed_b200 import engine as E
    assert E.padded_width(152) == 160 and E.padded_width(368) == 368 and E.padded_width(56) == 56 and E.padded_width(24) == 24
    assert E.fold_dim(152) == E.fold_dim(160)
    cfg = O.Config(feature_arch='rny002_gsf', clip_len=4, n_layers=2, sgp_ks=5, sgp_r=2, num_classes=4, radi_displacement=1, crop_dim=None)
    sd = {k: v.clone() for k, v in O.random_state(cfg, 3).items()}
    ref = {k: v.clone() for k, v in sd.items()}
    p = '_features.s3.b2'
    E.pad_block_state(sd, p, 152, 152, 160, 160, True)
    w1 = sd[p + '.conv1.net.conv.weight']
    assert tuple(w1.shape) == (160, 160, 1, 1) and torch.equal(w1[:152, :152], ref[p + '.conv1.net.conv.weight'])
    assert float(w1[152:].abs().max()) == 0 and float(w1[:, 152:].abs().max()) == 0
    for q in ('.conv1.net.bn', '.conv2.bn', '.conv3.bn'):
        assert float(sd[p + q + '.weight'][152:].abs().max()) == 0 and float(sd[p + q + '.bias'][152:].abs().max()) == 0
        assert float(sd[p + q + '.running_mean'][152:].abs().max()) == 0 and bool((sd[p + q + '.running_var'][152:] == 1).all())
    assert tuple(sd[p + '.conv2.conv.weight'].shape) == (160, 8, 3, 3) and float(sd[p + '.conv2.conv.weight'][152:].abs().max()) == 0
    assert tuple(sd[p + '.se.fc1.weight'].shape)[1] == 160 and tuple(sd[p + '.se.fc2.weight'].shape)[0] == 160
    assert tuple(sd[p + '.conv3.conv.weight'].shape) == (160, 160, 1, 1)
    # the gate-shift parameters of the block are untouched (the fold width is computed from the real width)
    for k in ref:
        if k.startswith(p + '.conv1.gs.'):
            assert torch.equal(sd[k], ref[k])
    # function check with plain torch ops on the padded vs original tensors: conv1 -> bn -> relu of a zero-padded input
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 152, 5, 4, generator=g)
    xp = torch.cat([x, torch.zeros(2, 8, 5, 4)], 1)

    def conv_bn_relu(inp, s, q):
        y = torch.nn.functional.conv2d(inp, s[q + '.conv.weight'])
        y = torch.nn.functional.batch_norm(y, s[q + '.bn.running_mean'], s[q + '.bn.running_var'], s[q + '.bn.weight'], s[q + '.bn.bias'], False, 0.0, 1e-5)
        return torch.relu(y)

    a, b = conv_bn_relu(x, ref, p + '.conv1.net'), conv_bn_relu(xp, sd, p + '.conv1.net')
    assert torch.equal(b[:, :152], a) and float(b[:, 152:].abs().max()) == 0
