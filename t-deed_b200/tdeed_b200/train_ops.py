"""Tensor-level wrappers over the training entry points (include/tdeed_b200_train.h).

Same contract as ops.py: PyTorch supplies device memory and the current stream; every computation happens inside
libtdeed_sm100.so.  Workspaces are allocated here per call (the caching allocator makes that cheap and stream-safe).
"""
import ctypes

import torch

from . import _lib as L


def _f32(n, dev):
    return torch.empty(max(int(n), 1), dtype=torch.float32, device=dev)


# ---- BatchNorm (training mode) ----

def bn_stats(x2d, C, gamma, beta, running_mean=None, running_var=None, eps=1e-5, momentum=0.1):
    """x2d: [M, ld] view whose first C columns are normalised.  -> stats fp32 [4, C] (mean, invstd, scale, shift)."""
    M, ld = x2d.shape[0], x2d.stride(0)
    stats = torch.empty((4, C), dtype=torch.float32, device=x2d.device)
    ws = _f32(L.load().tdeed_bn_workspace_floats(C), x2d.device)
    L.check(L.load().tdeed_bn_stats(L.dtype_code(x2d.dtype), L.ptr(x2d), M, C, ld, L.ptr(gamma), L.ptr(beta), eps, momentum,
                                    L.ptr(running_mean), L.ptr(running_var), L.ptr(stats), L.ptr(ws), L.stream()), 'bn_stats')
    return stats


def bn_act_fwd(y, stats, residual=None, relu=True, out=None):
    C = y.shape[-1]
    M = y.numel() // C
    if out is None:
        out = torch.empty_like(y)
    L.check(L.load().tdeed_bn_act_fwd(L.dtype_code(y.dtype), L.ptr(y), M, C, L.ptr(stats), L.ptr(residual), int(relu),
                                      L.ptr(out), L.stream()), 'bn_act_fwd')
    return out


def bn_act_bwd(dz, z, y, stats, want_dres=False, dy=None):
    """-> (dy, dgamma, dbeta, dres|None).  z=None: no ReLU."""
    C = y.shape[-1]
    M = y.numel() // C
    dev = y.device
    if dy is None:
        dy = torch.empty_like(y)
    dres = torch.empty_like(y) if want_dres else None
    dgamma = torch.empty(C, dtype=torch.float32, device=dev)
    dbeta = torch.empty(C, dtype=torch.float32, device=dev)
    ws = _f32(L.load().tdeed_bn_workspace_floats(C), dev)
    L.check(L.load().tdeed_bn_act_bwd(L.dtype_code(y.dtype), L.ptr(dz), L.ptr(z), L.ptr(y), M, C, L.ptr(stats), L.ptr(dgamma),
                                      L.ptr(dbeta), L.ptr(dy), L.ptr(dres), L.ptr(ws), L.stream()), 'bn_act_bwd')
    return dy, dgamma, dbeta, dres


# ---- GEMM-shaped gradients ----

def gemm_tn(a, b, m, n, rows, lda=None, ldb=None, gather=None, alpha=1.0, out=None):
    """out[m, n] = alpha * sum_r a[r, :m]^T b[r, :n]."""
    dev = a.device
    lda = lda if lda is not None else a.stride(-2)
    ldb = ldb if ldb is not None else b.stride(-2)
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=dev)
    ws = _f32(L.load().tdeed_gemm_tn_workspace_floats(rows, m, n), dev)
    gs, gh, gw = gather if gather else (1, 0, 0)
    L.check(L.load().tdeed_gemm_tn(L.dtype_code(a.dtype), L.ptr(a), lda, L.dtype_code(b.dtype), L.ptr(b), ldb, rows, m, n,
                                   gs, gh, gw, alpha, L.ptr(out), out.stride(0), L.ptr(ws), L.stream()), 'gemm_tn')
    return out


def colsum(x2d, C=None, out=None):
    M, ld = x2d.shape[0], x2d.stride(0)
    C = C if C is not None else x2d.shape[1]
    if out is None:
        out = torch.empty(C, dtype=torch.float32, device=x2d.device)
    ws = _f32(L.load().tdeed_colsum_workspace_floats(M, C), x2d.device)
    L.check(L.load().tdeed_colsum(L.dtype_code(x2d.dtype), L.ptr(x2d), M, C, ld, L.ptr(out), L.ptr(ws), L.stream()), 'colsum')
    return out


def strided_gather(src, stride):
    n, h, w, c = src.shape
    dst = torch.empty((n, (h + stride - 1) // stride, (w + stride - 1) // stride, c), dtype=src.dtype, device=src.device)
    L.check(L.load().tdeed_strided_gather(L.dtype_code(src.dtype), L.ptr(src), L.ptr(dst), n, h, w, c, stride, L.stream()), 'strided_gather')
    return dst


def stem_bwd_weight_tc(frames, unit_input, crop, flip, dy, out):
    """bf16 path: im2col of the normalised input (bf16 [P, 32]) + the tcgen05 dW GEMM; out fp32 [32, 3, 3, 3]."""
    n, _, in_h, in_w = frames.shape
    cy, cx, h, w = crop
    P_ = n * ((h + 1) // 2) * ((w + 1) // 2)
    patches = torch.empty((P_, 32), dtype=torch.bfloat16, device=frames.device)
    L.check(L.load().tdeed_stem_im2col(L.ptr(frames), L.dtype_code(frames.dtype), int(unit_input), n, in_h, in_w, cy, cx, h, w,
                                       int(bool(flip)), L.ptr(patches), L.stream()), 'stem_im2col')
    dw = gemm_tn(dy.view(P_, 32), patches, 32, 32, P_)
    out.view(32, 27).copy_(dw[:, :27])
    return out


def strided_add_(dst, src, stride):
    n, h, w, c = dst.shape
    L.check(L.load().tdeed_strided_add(L.dtype_code(dst.dtype), L.ptr(dst), L.ptr(src), n, h, w, c, stride, L.stream()),
            'strided_add')
    return dst


# ---- spatial convolutions ----

def stem_raw(frames, unit_input, crop, flip, weight, out_dtype):
    n, _, in_h, in_w = frames.shape
    cy, cx, h, w = crop
    out = torch.empty((n, (h + 1) // 2, (w + 1) // 2, 32), dtype=out_dtype, device=frames.device)
    L.check(L.load().tdeed_stem_raw_fwd(L.ptr(frames), L.dtype_code(frames.dtype), int(unit_input), n, in_h, in_w, cy, cx, h, w,
                                        int(bool(flip)), L.ptr(weight), L.ptr(out), L.dtype_code(out_dtype), L.stream()), 'stem_raw')
    return out


def stem_bwd_weight(frames, unit_input, crop, flip, dy, out=None):
    n, _, in_h, in_w = frames.shape
    cy, cx, h, w = crop
    if out is None:
        out = torch.empty((32, 3, 3, 3), dtype=torch.float32, device=frames.device)
    ws = _f32(L.load().tdeed_stem_bwd_weight_workspace_floats(), frames.device)
    L.check(L.load().tdeed_stem_bwd_weight(L.ptr(frames), L.dtype_code(frames.dtype), int(unit_input), n, in_h, in_w, cy, cx, h, w,
                                           int(bool(flip)), L.ptr(dy), L.dtype_code(dy.dtype), L.ptr(out), L.ptr(ws), L.stream()),
            'stem_bwd_weight')
    return out


def conv3x3g_raw(x, weight, group_width, stride):
    n, h, w, c = x.shape
    out = torch.empty((n, (h + stride - 1) // stride, (w + stride - 1) // stride, c), dtype=x.dtype, device=x.device)
    L.check(L.load().tdeed_conv3x3g_raw_fwd(L.dtype_code(x.dtype), L.ptr(x), n, h, w, c, group_width, stride, L.ptr(weight),
                                            L.ptr(out), L.stream()), 'conv3x3g_raw')
    return out


def conv3_weight_image(weight, group_width, transpose_flip=False):
    """fp32 [C, gw, 3, 3] -> bf16 UMMA B tiles for the tcgen05 grouped conv (transpose_flip: the data-gradient kernel)."""
    c = weight.shape[0]
    img = torch.empty(int(L.load().tdeed_conv3_weight_image_elems(c)), dtype=torch.bfloat16, device=weight.device)
    L.check(L.load().tdeed_conv3_weight_image(L.ptr(weight), c, group_width, int(transpose_flip), L.ptr(img), L.stream()),
            'conv3_weight_image')
    return img


def conv3x3g_tc_raw(x, wimg, stride):
    """bf16 NHWC raw grouped 3x3 convolution on tcgen05 (no bias, no activation)."""
    n, h, w, c = x.shape
    out = torch.empty((n, (h + stride - 1) // stride, (w + stride - 1) // stride, c), dtype=x.dtype, device=x.device)
    L.check(L.load().tdeed_conv3x3g_tc_raw_fwd(L.ptr(x), n, h, w, c, stride, L.ptr(wimg), L.ptr(out), L.stream()), 'conv3x3g_tc_raw')
    return out


def conv3x3g_tc_bwd_data_s2(dy, in_shape, weight, group_width):
    """bf16 data gradient of the stride-2 grouped conv on tcgen05 (four parity convolutions over the dy grid)."""
    n, h, w, c = in_shape
    elems = int(L.load().tdeed_conv3_weight_image_elems(c))
    imgs = torch.empty(4 * elems, dtype=torch.bfloat16, device=dy.device)
    for q in range(4):
        L.check(L.load().tdeed_conv3_weight_image(L.ptr(weight), c, group_width, 2 + q, imgs.data_ptr() + 2 * q * elems, L.stream()),
                'conv3_weight_image')
    dx = torch.empty(in_shape, dtype=dy.dtype, device=dy.device)
    L.check(L.load().tdeed_conv3x3g_tc_bwd_data_s2(L.ptr(dy), n, h, w, c, L.ptr(imgs), L.ptr(dx), L.stream()), 'conv3x3g_tc_bwd_data_s2')
    return dx


def conv3x3g_bwd_data(dy, in_shape, weight, group_width, stride):
    n, h, w, c = in_shape
    dx = torch.empty(in_shape, dtype=dy.dtype, device=dy.device)
    L.check(L.load().tdeed_conv3x3g_bwd_data(L.dtype_code(dy.dtype), L.ptr(dy), n, h, w, c, group_width, stride, L.ptr(weight),
                                             L.ptr(dx), L.stream()), 'conv3x3g_bwd_data')
    return dx


def conv3x3g_bwd_weight(x, dy, group_width, stride, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((c, group_width, 3, 3), dtype=torch.float32, device=x.device)
    ws = _f32(L.load().tdeed_conv3x3g_bwd_weight_workspace_floats(n, h, w, c, group_width, stride), x.device)
    L.check(L.load().tdeed_conv3x3g_bwd_weight(L.dtype_code(x.dtype), L.ptr(x), L.ptr(dy), n, h, w, c, group_width, stride,
                                               L.ptr(out), L.ptr(ws), L.stream()), 'conv3x3g_bwd_weight')
    return out


# ---- squeeze-excite / pool ----

def se_train_fwd(x, w1, b1, w2t, b2):
    """-> (out, fwd_workspace[mean | scale])."""
    n, h, w, c = x.shape
    out = torch.empty_like(x)
    ws = _f32(L.load().tdeed_se_workspace_floats(n, c), x.device)
    L.check(L.load().tdeed_se_train_fwd(L.dtype_code(x.dtype), L.ptr(x), L.ptr(out), n, h * w, c, w1.shape[0], L.ptr(w1), L.ptr(b1),
                                        L.ptr(w2t), L.ptr(b2), L.ptr(ws), L.stream()), 'se_train_fwd')
    return out, ws


def se_bwd(x, du, w1, b1, w2t, fwd_ws):
    """-> (dx, d_fc1_w [rd,c], d_fc1_b, d_fc2_w [c,rd], d_fc2_b)."""
    n, h, w, c = x.shape
    rd = w1.shape[0]
    dx = torch.empty_like(x)
    vec = _f32(L.load().tdeed_se_bwd_vec_floats(n, c, rd), x.device)
    L.check(L.load().tdeed_se_bwd(L.dtype_code(x.dtype), L.ptr(x), L.ptr(du), n, h * w, c, rd, L.ptr(w1), L.ptr(b1), L.ptr(w2t),
                                  L.ptr(fwd_ws), L.ptr(dx), L.ptr(vec), L.stream()), 'se_bwd')
    dv = vec[:n * c].view(n, c)
    dh = vec[n * c:n * c + n * rd].view(n, rd)
    hh = vec[n * c + n * rd:n * c + 2 * n * rd].view(n, rd)
    mean = fwd_ws[:n * c].view(n, c)
    d_w2 = gemm_tn(dv, hh, c, rd, n)
    d_b2 = colsum(dv)
    d_w1 = gemm_tn(dh, mean, rd, c, n)
    d_b1 = colsum(dh)
    return dx, d_w1, d_b1, d_w2, d_b2


def pool_posenc_bwd(dfeat, clips, clip_len, hw, c, dtype):
    n = clips * clip_len
    dz = torch.empty((n, hw, c), dtype=dtype, device=dfeat.device)
    dte = torch.empty((clip_len, c), dtype=torch.float32, device=dfeat.device)
    L.check(L.load().tdeed_pool_posenc_bwd(L.dtype_code(dtype), L.ptr(dfeat), clips, clip_len, hw, c, L.ptr(dz), L.ptr(dte),
                                           L.stream()), 'pool_posenc_bwd')
    return dz, dte


# ---- gate-shift ----

def gsf_cat_fwd(x, clips, clip_len, fold, mode, stats, conv3d_w, conv3d_b, cc_w, cc_b):
    """x NHWC (clips*clip_len, h, w, c).  -> (cat [N*h*w, c], fwd workspace)."""
    n, h, w, c = x.shape
    ws = _f32(L.load().tdeed_gsf_workspace_floats(clips, clip_len, h, w, fold), x.device)
    out = torch.empty((n * h * w, c), dtype=x.dtype, device=x.device)
    L.check(L.load().tdeed_gsf_cat_fwd(L.dtype_code(x.dtype), mode, L.ptr(x), clips, clip_len, h, w, c, fold, L.ptr(stats[2]),
                                       L.ptr(stats[3]), L.ptr(conv3d_w), L.ptr(conv3d_b), L.ptr(cc_w), L.ptr(cc_b), L.ptr(ws),
                                       L.ptr(out), L.stream()), 'gsf_cat_fwd')
    return out, ws


def gsf_bwd(x, dcat, add, clips, clip_len, fold, mode, stats, conv3d_w, cc_w, fwd_ws):
    """-> (dx like x, d_conv3d_w [fold*27], d_conv3d_b [2], d_cc [2,19] | None, d_gamma [fold], d_beta [fold])."""
    n, h, w, c = x.shape
    dev = x.device
    ws = _f32(L.load().tdeed_gsf_bwd_workspace_floats(clips, clip_len, h, w, fold), dev)
    dx = torch.empty_like(x)
    dw3 = torch.empty(fold * 27, dtype=torch.float32, device=dev)
    db3 = torch.empty(2, dtype=torch.float32, device=dev)
    dcc = torch.empty((2, 19), dtype=torch.float32, device=dev) if mode == L.SHIFT_GSF else None
    dg = torch.empty(fold, dtype=torch.float32, device=dev)
    db = torch.empty(fold, dtype=torch.float32, device=dev)
    L.check(L.load().tdeed_gsf_bwd(L.dtype_code(x.dtype), mode, L.ptr(x), L.ptr(dcat), L.ptr(add), clips, clip_len, h, w, c, fold,
                                   L.ptr(stats), L.ptr(conv3d_w), L.ptr(cc_w), L.ptr(fwd_ws), L.ptr(ws), L.ptr(dx), L.ptr(dw3),
                                   L.ptr(db3), L.ptr(dcc), L.ptr(dg), L.ptr(db), L.stream()), 'gsf_bwd')
    return dx, dw3, db3, dcc, dg, db


# ---- temporal layers ----

def chan_ln_fwd(x, T, w, b, want_pool=False):
    """x [B, t_in, C] fp32 -> (ln [B,T,C], stats [B*T,2], xp|None, argmax|None)."""
    B, t_in, C = x.shape
    dev = x.device
    ln = torch.empty((B, T, C), dtype=torch.float32, device=dev)
    stats = torch.empty((B * T, 2), dtype=torch.float32, device=dev)
    xp = torch.empty((B, T, C), dtype=torch.float32, device=dev) if want_pool else None
    arg = torch.empty((B, T, C), dtype=torch.int32, device=dev) if want_pool else None
    L.check(L.load().tdeed_chan_ln_fwd(L.ptr(x), B, t_in, T, C, L.ptr(w), L.ptr(b), L.ptr(xp), L.ptr(arg), L.ptr(ln), L.ptr(stats),
                                       L.stream()), 'chan_ln_fwd')
    return ln, stats, xp, arg


def chan_ln_bwd(xp, stats, dln, ld, w, add=None):
    """xp [rows, C] (flattened ok), dln with leading dim ld.  -> (dx, dw, db)."""
    C = xp.shape[-1]
    rows = xp.numel() // C
    dev = xp.device
    dx = torch.empty_like(xp)
    dw = torch.empty(C, dtype=torch.float32, device=dev)
    db = torch.empty(C, dtype=torch.float32, device=dev)
    L.check(L.load().tdeed_chan_ln_bwd(L.ptr(xp), L.ptr(stats), L.ptr(dln), ld, rows, C, L.ptr(w), L.ptr(add), L.ptr(dx), L.ptr(dw),
                                       L.ptr(db), L.stream()), 'chan_ln_bwd')
    return dx, dw, db


def maxpool_bwd(dxp, argmax, t_in, add=None):
    B, T, C = dxp.shape
    dx = torch.empty((B, t_in, C), dtype=torch.float32, device=dxp.device)
    L.check(L.load().tdeed_maxpool_bwd(L.ptr(dxp), L.ptr(argmax), B, t_in, T, C, L.ptr(add), L.ptr(dx), L.stream()), 'maxpool_bwd')
    return dx


BRANCH_ORDER = ('psi_w', 'psi_b', 'convw_w', 'convw_b', 'convkw_w', 'convkw_b', 'fc_w', 'fc_b', 'gfc_w', 'gfc_b')


def sgp_branch_bwd(ln, ld_ln, d_conv, d_fc, d_id, ld_g, B, T, C, ks, up, weights, grads):
    """weights / grads: dicts keyed by BRANCH_ORDER (fp32 device tensors; grads are written).  -> d_ln [B,T,C]."""
    dev = weights['psi_w'].device
    wa = (L.c_vp * 10)(*[L.ptr(weights[k]) for k in BRANCH_ORDER])
    ga = (L.c_vp * 10)(*[L.ptr(grads[k]) for k in BRANCH_ORDER])
    d_ln = torch.empty((B, T, C), dtype=torch.float32, device=dev)
    ws = _f32(L.load().tdeed_sgp_branch_bwd_workspace_floats(B, T, C), dev)
    L.check(L.load().tdeed_sgp_branch_bwd(L.ptr(ln), ld_ln, L.ptr(d_conv), L.ptr(d_fc), L.ptr(d_id), ld_g, B, T, C, ks, up,
                                          wa, ga, L.ptr(d_ln), L.ptr(ws), L.stream()), 'sgp_branch_bwd')
    return d_ln


def groupnorm_bwd(y, dg, gamma, add=None, groups=16):
    B, T, C = y.shape
    dev = y.device
    dy = torch.empty_like(y)
    dgamma = torch.empty(C, dtype=torch.float32, device=dev)
    dbeta = torch.empty(C, dtype=torch.float32, device=dev)
    ws = _f32(2 * B * C, dev)
    L.check(L.load().tdeed_groupnorm_bwd(L.ptr(y), L.ptr(dg), B, T, C, groups, L.ptr(gamma), L.ptr(add), L.ptr(dy), L.ptr(dgamma),
                                         L.ptr(dbeta), L.ptr(ws), L.stream()), 'groupnorm_bwd')
    return dy, dgamma, dbeta


def gelu_fwd(h, out_dtype):
    out = torch.empty(h.shape, dtype=out_dtype, device=h.device)
    L.check(L.load().tdeed_gelu_fwd(L.ptr(h), h.numel(), L.ptr(out), L.dtype_code(out_dtype), L.stream()), 'gelu_fwd')
    return out


def gelu_bwd(h, da, out_dtype):
    dh = torch.empty(h.shape, dtype=out_dtype, device=h.device)
    L.check(L.load().tdeed_gelu_bwd(L.ptr(h), L.ptr(da), h.numel(), L.ptr(dh), L.dtype_code(out_dtype), L.stream()), 'gelu_bwd')
    return dh


def upsample_bwd(dxu, t_coarse):
    B, T, C = dxu.shape
    dx = torch.empty((B, t_coarse, C), dtype=torch.float32, device=dxu.device)
    L.check(L.load().tdeed_upsample_bwd(L.ptr(dxu), B, t_coarse, T, C, L.ptr(dx), L.stream()), 'upsample_bwd')
    return dx


def cast(x, out_dtype, out=None):
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    L.check(L.load().tdeed_cast_f32(L.ptr(x), x.numel(), L.ptr(out), L.dtype_code(out_dtype), L.stream()), 'cast')
    return out


# ---- heads / loss / optimizer ----

def dropout_fwd(x, p, seed):
    out = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    L.check(L.load().tdeed_dropout_fwd(L.ptr(x), x.numel(), p, seed, L.ptr(out), L.ptr(mask), L.stream()), 'dropout_fwd')
    return out, mask


def dropout_fwd_devseed(x, p, seed_dev, salt):
    """seed_dev: int64 device tensor [1] (the step counter); mask = f(seed_dev[0]*2 + salt, element index)."""
    out = torch.empty_like(x)
    mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    L.check(L.load().tdeed_dropout_fwd_devseed(L.ptr(x), x.numel(), p, L.ptr(seed_dev), salt, L.ptr(out), L.ptr(mask), L.stream()),
            'dropout_fwd_devseed')
    return out, mask


def dropout_bwd(dy, mask, p, add=None):
    dx = torch.empty_like(dy)
    L.check(L.load().tdeed_dropout_bwd(L.ptr(dy), L.ptr(mask), dy.numel(), p, L.ptr(add), L.ptr(dx), L.stream()), 'dropout_bwd')
    return dx


def linear_fwd(x2d, W, b, out=None, col0=0):
    M, C = x2d.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x2d.device)
    L.check(L.load().tdeed_linear_fwd(L.ptr(x2d), M, C, L.ptr(W), L.ptr(b), N, out.data_ptr() + 4 * col0, out.stride(0), L.stream()),
            'linear_fwd')
    return out


def linear_bwd_data(dout, W, add=None, col0=0, n=None):
    M = dout.shape[0]
    N, C = (n if n is not None else W.shape[0]), W.shape[1]
    dx = torch.empty((M, C), dtype=torch.float32, device=dout.device)
    L.check(L.load().tdeed_linear_bwd_data(dout.data_ptr() + 4 * col0, dout.stride(0), M, C, L.ptr(W), N, L.ptr(add), L.ptr(dx),
                                           L.stream()), 'linear_bwd_data')
    return dx


def mixup_u8(a, b, lam):
    """a, b: uint8 (n, ...) clip batches on the device, lam: fp32 (n, 2) device tensor -> fp32 lam[:,0]*a + lam[:,1]*b (one pass,
    the reference's roundings: model/model.py:228-254)."""
    assert a.dtype == torch.uint8 and b.dtype == torch.uint8 and a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    assert lam.dtype == torch.float32 and lam.is_contiguous() and tuple(lam.shape) == (a.shape[0], 2)
    out = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    L.check(L.load().tdeed_mixup_u8(L.ptr(a), L.ptr(b), L.ptr(lam), a.shape[0], a[0].numel(), L.ptr(out), L.stream()), 'mixup_u8')
    return out


def ce_mse_loss(logits, target_hard, target_soft, class_weight, displ, labelD):
    """logits [M, K] fp32.  -> (loss[3] device tensor, dlogits [M,K], ddispl [M] | None)."""
    M, K = logits.shape
    dev = logits.device
    loss = torch.empty(3, dtype=torch.float32, device=dev)
    dlogits = torch.empty((M, K), dtype=torch.float32, device=dev)
    ddispl = torch.empty(M, dtype=torch.float32, device=dev) if displ is not None else None
    L.check(L.load().tdeed_ce_mse_loss(L.ptr(logits), M, K, logits.stride(0), L.ptr(target_hard), L.ptr(target_soft),
                                       L.ptr(class_weight), L.ptr(displ), L.ptr(labelD), L.ptr(loss), L.ptr(dlogits), L.ptr(ddispl),
                                       L.stream()), 'ce_mse_loss')
    return loss, dlogits, ddispl


def ce_mse_loss_2heads(logits, B, T, n1, n2, dataset, target_hard, target_soft, class_weight, displ, labelD):
    """logits [B*T, n1+n2] fp32; dataset int32 [B].  -> (loss[3], dlogits [B*T, n1+n2], ddispl | None)."""
    dev = logits.device
    loss = torch.empty(3, dtype=torch.float32, device=dev)
    dlogits = torch.empty((B * T, n1 + n2), dtype=torch.float32, device=dev)
    ddispl = torch.empty(B * T, dtype=torch.float32, device=dev) if displ is not None else None
    L.check(L.load().tdeed_ce_mse_loss_2heads(L.ptr(logits), B, T, n1, n2, logits.stride(0), L.ptr(dataset), L.ptr(target_hard),
                                              L.ptr(target_soft), L.ptr(class_weight), L.ptr(displ), L.ptr(labelD), L.ptr(loss),
                                              L.ptr(dlogits), L.ptr(ddispl), L.stream()), 'ce_mse_loss_2heads')
    return loss, dlogits, ddispl


def adamw_step_(p, g, m, v, lr, betas, eps, weight_decay, step, grad_scale=1.0, shadow=None):
    L.check(L.load().tdeed_adamw_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, betas[0], betas[1], eps, weight_decay,
                                      step, grad_scale, L.ptr(shadow), L.stream()), 'adamw_step')


def axpy_(x, alpha, y):
    L.check(L.load().tdeed_axpy(L.ptr(x), alpha, x.numel(), L.ptr(y), L.stream()), 'axpy')
    return y
