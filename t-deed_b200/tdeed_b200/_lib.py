"""ctypes binding of libtdeed_sm100.so (C-ABI declared in include/tdeed_b200.h).

The library is the product: there is no PyTorch / CPU fallback.  If it is missing or an entry
point fails, a RuntimeError is raised with the library's own error text.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libtdeed_sm100.so')

F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
SHIFT_GSM, SHIFT_GSF = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05, GEMM_TCGEN05_THIN = 0, 1, 2, 3
GEMM_MAX_SEGS = 2
MAX_CLIPS_PER_CALL = 128
MAX_STARTS_PER_CALL = 256

c_int, c_ll, c_float, c_double, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_void_p


class GemmSeg(ctypes.Structure):
    _fields_ = [('a', c_vp), ('lda', c_ll), ('col0', c_int), ('k', c_int)]


class SgpWeights(ctypes.Structure):
    _names = ('ln_w', 'ln_b', 'gn_w', 'gn_b', 'psi_w', 'psi_b', 'fc_w', 'fc_b', 'convw_w', 'convw_b',
              'convkw_w', 'convkw_b', 'gfc_w', 'gfc_b')
    _fields_ = [(n, c_vp) for n in _names]


class MixerWeights(ctypes.Structure):
    _names = ('ln1_w', 'ln1_b', 'ln2_w', 'ln2_b', 'psi1_w', 'psi1_b', 'psi2_w', 'psi2_b', 'convw1_w', 'convw1_b',
              'convkw1_w', 'convkw1_b', 'convw2_w', 'convw2_b', 'convkw2_w', 'convkw2_b', 'fc1_w', 'fc1_b',
              'gfc1_w', 'gfc1_b', 'fc2_w', 'fc2_b', 'gfc2_w', 'gfc2_b')
    _fields_ = [(n, c_vp) for n in _names]


# name -> (restype, argtypes); mirrors include/tdeed_b200.h one to one
SIGNATURES = {
    'tdeed_abi_version': (c_int, []),
    'tdeed_last_error': (ctypes.c_char_p, []),
    'tdeed_stem_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp,
                               c_vp, c_int, c_vp]),
    'tdeed_stem_tc_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                                  c_vp, c_int, c_vp, c_int, c_vp, c_vp]),
    'tdeed_stem_tc2_wimg_bytes': (c_ll, []),
    'tdeed_stem_tc2_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, ctypes.POINTER(c_float), c_vp,
                                   c_vp, c_int, c_vp, c_int, c_vp, c_vp]),
    'tdeed_gemm_fwd': (c_int, [c_int, c_ll, c_int, c_int, ctypes.POINTER(GemmSeg), c_int, c_int, c_int, c_vp, c_vp,
                               c_vp, c_ll, c_int, c_int, c_vp, c_ll, c_int, c_int, c_vp]),
    'tdeed_conv3x3g_fwd': (c_int, [c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_conv3x3g_tc_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_se_workspace_floats': (c_ll, [c_int, c_int]),
    'tdeed_se_fwd': (c_int, [c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_se_gate_fwd': (c_int, [c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_gemm_scaled_fwd': (c_int, [c_ll, c_int, c_int, c_vp, c_ll, c_vp, c_int, c_vp, c_vp, c_vp, c_ll, c_int, c_vp, c_ll, c_vp]),
    'tdeed_gsf_workspace_floats': (c_ll, [c_int, c_int, c_int, c_int, c_int]),
    'tdeed_gsf_fwd': (c_int, [c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                              c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    'tdeed_gsf_fwd_natural': (c_int, [c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    'tdeed_gsf_interleaved_position': (c_int, [c_int, c_int]),
    'tdeed_pool_posenc_fwd': (c_int, [c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_sgp_mix_workspace_floats': (c_ll, [c_int, c_int, c_int]),
    'tdeed_sgp_mix_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(SgpWeights), c_vp, c_vp,
                                  c_vp, c_int, c_vp]),
    'tdeed_sgp_mixer_workspace_floats': (c_ll, [c_int, c_int, c_int, c_int]),
    'tdeed_sgp_mixer_mix_fwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                        ctypes.POINTER(MixerWeights), c_vp, c_vp, c_int, c_vp]),
    'tdeed_groupnorm_workspace_floats': (c_ll, [c_int, c_int, c_int, c_int]),
    'tdeed_groupnorm_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    'tdeed_heads_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp,
                                c_vp]),
    'tdeed_softmax_scatter_fwd': (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    'tdeed_clip_accumulate': (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    'tdeed_extract_events': (c_int, [c_vp, c_vp, c_int, c_int, c_float, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp]),
    'tdeed_nms_workspace_bytes': (c_ll, [c_int, c_int]),
    'tdeed_nms': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_double, c_int, c_vp, c_vp, c_vp, c_vp,
                          c_vp, c_vp]),
    'tdeed_gather_rows': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_ll, c_vp]),
    'tdeed_scatter_rows_ring': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_ll, c_vp]),
    'tdeed_gather_clip_rows': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                       ctypes.POINTER(c_int), c_int, c_ll, c_vp]),
    'tdeed_clip_accumulate_host': (c_int, [c_vp, c_vp, c_int, c_int, c_vp, ctypes.POINTER(c_int), c_int, c_int, c_int, c_vp]),
    'tdeed_match_events': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
}

# include/tdeed_b200_train.h
c_ull = ctypes.c_ulonglong
SIGNATURES.update({
    'tdeed_bn_workspace_floats': (c_ll, [c_int]),
    'tdeed_bn_stats': (c_int, [c_int, c_vp, c_ll, c_int, c_ll, c_vp, c_vp, c_float, c_float, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_bn_act_fwd': (c_int, [c_int, c_vp, c_ll, c_int, c_vp, c_vp, c_int, c_vp, c_vp]),
    'tdeed_bn_act_bwd': (c_int, [c_int, c_vp, c_vp, c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_gemm_tn_workspace_floats': (c_ll, [c_ll, c_int, c_int]),
    'tdeed_gemm_tn': (c_int, [c_int, c_vp, c_ll, c_int, c_vp, c_ll, c_ll, c_int, c_int, c_int, c_int, c_int, c_float, c_vp,
                              c_ll, c_vp, c_vp]),
    'tdeed_colsum_workspace_floats': (c_ll, [c_ll, c_int]),
    'tdeed_colsum': (c_int, [c_int, c_vp, c_ll, c_int, c_ll, c_vp, c_vp, c_vp]),
    'tdeed_strided_gather': (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'tdeed_stem_im2col': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'tdeed_strided_add': (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    'tdeed_stem_raw_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp,
                                   c_int, c_vp]),
    'tdeed_stem_bwd_weight_workspace_floats': (c_ll, []),
    'tdeed_stem_bwd_weight': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int,
                                      c_vp, c_vp, c_vp]),
    'tdeed_conv3x3g_raw_fwd': (c_int, [c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_conv3_weight_image_elems': (c_ll, [c_int]),
    'tdeed_conv3_weight_image': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp]),
    'tdeed_conv3x3g_tc_raw_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_conv3x3g_tc_bwd_data_s2': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_conv3x3g_bwd_data': (c_int, [c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_conv3x3g_bwd_weight_workspace_floats': (c_ll, [c_int, c_int, c_int, c_int, c_int, c_int]),
    'tdeed_conv3x3g_bwd_weight': (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_se_train_fwd': (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_se_bwd_vec_floats': (c_ll, [c_int, c_int, c_int]),
    'tdeed_se_bwd': (c_int, [c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_pool_posenc_bwd': (c_int, [c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_gsf_cat_fwd': (c_int, [c_int, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                  c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_gsf_bwd_workspace_floats': (c_ll, [c_int, c_int, c_int, c_int, c_int]),
    'tdeed_gsf_bwd': (c_int, [c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp,
                              c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_chan_ln_fwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_chan_ln_bwd': (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_maxpool_bwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'tdeed_sgp_branch_bwd_workspace_floats': (c_ll, [c_int, c_int, c_int]),
    'tdeed_sgp_branch_bwd': (c_int, [c_vp, c_ll, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int,
                                     ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), c_vp, c_vp, c_vp]),
    'tdeed_groupnorm_bwd': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_gelu_fwd': (c_int, [c_vp, c_ll, c_vp, c_int, c_vp]),
    'tdeed_gelu_bwd': (c_int, [c_vp, c_vp, c_ll, c_vp, c_int, c_vp]),
    'tdeed_upsample_bwd': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'tdeed_cast_f32': (c_int, [c_vp, c_ll, c_vp, c_int, c_vp]),
    'tdeed_aug_color': (c_int, [c_vp, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_float,
                                c_int, c_float, c_vp, c_vp]),
    'tdeed_mixup_u8': (c_int, [c_vp, c_vp, c_vp, c_int, c_ll, c_vp, c_vp]),
    'tdeed_aug_gray_mean': (c_int, [c_vp, c_int, c_int, c_vp, c_vp]),
    'tdeed_aug_contrast_blur_flip': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_float, c_vp, c_int, ctypes.POINTER(c_float), c_int,
                                             c_vp, c_vp]),
    'tdeed_dropout_fwd': (c_int, [c_vp, c_ll, c_float, c_ull, c_vp, c_vp, c_vp]),
    'tdeed_dropout_fwd_devseed': (c_int, [c_vp, c_ll, c_float, c_vp, c_ull, c_vp, c_vp, c_vp]),
    'tdeed_dropout_bwd': (c_int, [c_vp, c_vp, c_ll, c_float, c_vp, c_vp, c_vp]),
    'tdeed_linear_fwd': (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_vp]),
    'tdeed_linear_bwd_data': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    'tdeed_ce_mse_loss': (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'tdeed_ce_mse_loss_2heads': (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                         c_vp]),
    'tdeed_adamw_step': (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_double, c_double, c_double, c_double, c_double, c_int,
                                 c_float, c_vp, c_vp]),
    'tdeed_axpy': (c_int, [c_vp, c_float, c_ll, c_vp, c_vp]),
})

_lib = None


def load():
    """dlopen the library (once) and declare every prototype.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('libtdeed_sm100.so not found at %s — build it with `python t-deed_b200/build.py` '
                           '(there is no CPU / PyTorch fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().tdeed_last_error().decode('utf-8', 'replace')
        raise RuntimeError('libtdeed_sm100 %s failed (%d): %s' % (what, rc, msg))


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.uint8:
        return U8
    raise TypeError('unsupported dtype %s' % dt)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, 'libtdeed_sm100 only takes CUDA tensors (no CPU path)'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
