"""Tensor-level wrappers over the C-ABI (one function per entry point of include/tdeed_b200.h).

PyTorch supplies device memory and the current stream only; every computation happens inside
libtdeed_sm100.so.  All wrappers are asynchronous on torch's current CUDA stream.
"""
import ctypes

import torch

from . import _lib as L


def stem(frames, crop, flip, weight, bias, out_dtype):
    """frames (N,3,H,W) u8|f32 (device) -> NHWC (N, ceil(h/2), ceil(w/2), 32)."""
    n, _, in_h, in_w = frames.shape
    cy, cx, h, w = crop
    out = torch.empty((n, (h + 1) // 2, (w + 1) // 2, 32), dtype=out_dtype, device=frames.device)
    L.check(L.load().tdeed_stem_fwd(L.ptr(frames), L.dtype_code(frames.dtype), n, in_h, in_w, cy, cx, h, w,
                                    int(bool(flip)), L.ptr(weight), L.ptr(bias), L.ptr(out), L.dtype_code(out_dtype),
                                    L.stream()), 'stem')
    return out


def stem_tc(frames, crop, flip, w0, b0, w1=None, b1=None, n1=0, want_stem=True, stem_sub=1):
    """bf16 tcgen05 stem (+ fused s1.b1.conv1).  Returns (stem_out | None, conv1_out | None), both NHWC bf16."""
    n, _, in_h, in_w = frames.shape
    cy, cx, h, w = crop
    oh, ow = (h + 1) // 2, (w + 1) // 2
    dev = frames.device
    out_stem = torch.empty((n, (oh + stem_sub - 1) // stem_sub, (ow + stem_sub - 1) // stem_sub, 32),
                           dtype=torch.bfloat16, device=dev) if want_stem else None
    out_c1 = torch.empty((n, oh, ow, n1), dtype=torch.bfloat16, device=dev) if w1 is not None else None
    L.check(L.load().tdeed_stem_tc_fwd(L.ptr(frames), L.dtype_code(frames.dtype), n, in_h, in_w, cy, cx, h, w,
                                       int(bool(flip)), L.ptr(w0), L.ptr(b0), L.ptr(w1), L.ptr(b1), n1,
                                       L.ptr(out_stem), stem_sub, L.ptr(out_c1), L.stream()), 'stem_tc')
    return out_stem, out_c1


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
STEM2_PAIRS = ((0, 2), (6, 8), (1, 7), (3, 5), (4, None))     # taps (ky*3+kx) sharing one K=16 MMA


def stem_tc2_weights(stem_w, stem_b):
    """Folded stem conv weights fp32 [32,3,3,3] / bias [32] (BN folded, for NORMALISED input) -> (wimg bf16, b0 fp32, pad_rgb) for
    tdeed_stem_tc2_fwd: weights for raw 0..255 input, bias absorbing the -mean/std term, raw padding pixel 255*mean."""
    dev = stem_w.device
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32, device=dev)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32, device=dev)
    w = (stem_w.float().reshape(32, 3, 9) / (255.0 * std.view(1, 3, 1))).to(torch.bfloat16)          # [co][ci][tap]
    pad = (255.0 * mean)
    b0 = stem_b.float() - (w.float() * pad.view(1, 3, 1)).sum(dim=(1, 2))
    tiles = torch.zeros((len(STEM2_PAIRS), 32, 16), dtype=torch.bfloat16, device=dev)
    for q, (ta, tb) in enumerate(STEM2_PAIRS):
        tiles[q, :, 0:3] = w[:, :, ta]
        if tb is not None:
            tiles[q, :, 8:11] = w[:, :, tb]
    img = tiles.reshape(len(STEM2_PAIRS), 4, 8, 2, 8).permute(0, 1, 3, 2, 4).contiguous()           # [q][n/8][k/8][n%8][k%8]
    return img.reshape(-1), b0.contiguous(), [float(v) for v in pad]


def stem_tc2(frames, crop, flip, wimg, b0, pad_rgb, w1, b1, n1, stem_sub=2):
    """uint8 frames (N,3,H,W) -> (stem_out subsample (N, ceil(oh/sub), ceil(ow/sub), 32), conv1_out (N, oh, ow, n1)), bf16."""
    n, _, in_h, in_w = frames.shape
    cy, cx, h, w = crop
    oh, ow = (h + 1) // 2, (w + 1) // 2
    dev = frames.device
    out_stem = torch.empty((n, (oh + stem_sub - 1) // stem_sub, (ow + stem_sub - 1) // stem_sub, 32), dtype=torch.bfloat16, device=dev)
    out_c1 = torch.empty((n, oh, ow, n1), dtype=torch.bfloat16, device=dev)
    pad = (ctypes.c_float * 3)(*pad_rgb)
    L.check(L.load().tdeed_stem_tc2_fwd(L.ptr(frames), n, in_h, in_w, cy, cx, h, w, int(bool(flip)), L.ptr(wimg), L.ptr(b0), pad,
                                        L.ptr(w1), L.ptr(b1), n1, L.ptr(out_stem), stem_sub, L.ptr(out_c1), L.stream()), 'stem_tc2')
    return out_stem, out_c1


def gemm(segs, weight, bias=None, residual=None, act=L.ACT_NONE, out=None, out_dtype=None, gather=None,
         backend=L.GEMM_AUTO, rows=None):
    """out[M,N] = act(concat_k(segs) @ weight.T + bias + residual).

    segs: list of (tensor2d_or_flat, lda, col0, k); tensors are activations whose rows are M.
    gather: (stride, h, w) for the strided 1x1 shortcut conv (rows = frames*ceil(h/s)*ceil(w/s)).
    """
    dtype = weight.dtype
    n_out = weight.shape[0]
    m = rows
    arr = (L.GemmSeg * len(segs))()
    for i, (t, lda, col0, k) in enumerate(segs):
        assert t.dtype == dtype, 'segment dtype %s != weight dtype %s' % (t.dtype, dtype)
        arr[i].a, arr[i].lda, arr[i].col0, arr[i].k = L.ptr(t), lda, col0, k
    if out is None:
        out = torch.empty((m, n_out), dtype=out_dtype or dtype, device=weight.device)
    ldo = out.stride(-2) if out.dim() >= 2 else n_out
    gs, gh, gw = gather if gather else (1, 0, 0)
    L.check(L.load().tdeed_gemm_fwd(L.dtype_code(dtype), m, n_out, len(segs), arr, gs, gh, gw, L.ptr(weight),
                                    L.ptr(bias), L.ptr(residual),
                                    residual.stride(-2) if residual is not None else 0,
                                    L.dtype_code(residual.dtype) if residual is not None else 0,
                                    act, L.ptr(out), ldo, L.dtype_code(out.dtype), backend, L.stream()), 'gemm')
    return out


def conv3x3g(x, weight, bias, group_width, stride, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, (h + stride - 1) // stride, (w + stride - 1) // stride, c), dtype=x.dtype, device=x.device)
    L.check(L.load().tdeed_conv3x3g_fwd(L.dtype_code(x.dtype), L.ptr(x), n, h, w, c, group_width, stride,
                                        L.ptr(weight), L.ptr(bias), L.ptr(out), L.stream()), 'conv3x3g')
    return out


def conv3_weight_image(w, group_width):
    """Folded conv2 weights fp32 [C][gw][3][3] -> bf16 UMMA B tiles [ceil(C/16)][9][16 out][16 in] in the canonical
    K-major no-swizzle layout ([n/8][k/8][n%8][k%8], 512 B per tile) consumed by tdeed_conv3x3g_tc_fwd."""
    c = w.shape[0]
    pairs = (c + 15) // 16
    co = torch.arange(c, device=w.device)
    tiles = torch.zeros((pairs, 9, 16, 16), dtype=torch.float32, device=w.device)
    for ci in range(group_width):
        k = (co // group_width) * group_width + ci - 16 * (co // 16)
        tiles[co // 16, :, co % 16, k] = w[:, ci].reshape(c, 9).float()
    img = tiles.reshape(pairs, 9, 2, 8, 2, 8).permute(0, 1, 2, 4, 3, 5).contiguous()
    return img.to(torch.bfloat16).reshape(-1)


def conv3x3g_tc(x, wimg, bias, stride, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, (h + stride - 1) // stride, (w + stride - 1) // stride, c), dtype=x.dtype, device=x.device)
    L.check(L.load().tdeed_conv3x3g_tc_fwd(L.ptr(x), n, h, w, c, stride, L.ptr(wimg), L.ptr(bias), L.ptr(out), L.stream()),
            'conv3x3g_tc')
    return out


def se_(x, w1, b1, w2, b2):
    n, h, w, c = x.shape
    ws = torch.empty(int(L.load().tdeed_se_workspace_floats(n, c)), dtype=torch.float32, device=x.device)
    L.check(L.load().tdeed_se_fwd(L.dtype_code(x.dtype), L.ptr(x), n, h * w, c, w1.shape[0], L.ptr(w1), L.ptr(b1),
                                  L.ptr(w2), L.ptr(b2), L.ptr(ws), L.stream()), 'se')
    return x


def se_gate(x, w1, b1, w2, b2):
    """Squeeze-excite gate of x NHWC (n,h,w,c): returns the fp32 gate [n, c] (x is NOT modified); the consumer applies it
    (gemm_scaled)."""
    n, h, w, c = x.shape
    ws = torch.empty(int(L.load().tdeed_se_workspace_floats(n, c)), dtype=torch.float32, device=x.device)
    L.check(L.load().tdeed_se_gate_fwd(L.dtype_code(x.dtype), L.ptr(x), n, h * w, c, w1.shape[0], L.ptr(w1), L.ptr(b1),
                                       L.ptr(w2), L.ptr(b2), L.ptr(ws), L.stream()), 'se_gate')
    return ws[n * c:].view(n, c)


def gemm_scaled(a, gate, rows_per_gate, weight, bias=None, residual=None, act=L.ACT_NONE, out=None):
    """out[M,N] = act((a * gate[row // rows_per_gate]) @ weight.T + bias + residual): conv3 with the SE gate folded into its
    A operand (bf16 tcgen05 kernel).  a: [M, K] bf16; gate: [M / rows_per_gate, K] fp32."""
    m, k = a.shape
    n_out = weight.shape[0]
    assert a.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and gate.dtype == torch.float32 and gate.is_contiguous()
    if out is None:
        out = torch.empty((m, n_out), dtype=torch.bfloat16, device=a.device)
    L.check(L.load().tdeed_gemm_scaled_fwd(m, n_out, k, L.ptr(a), a.stride(0), L.ptr(gate), rows_per_gate, L.ptr(weight), L.ptr(bias),
                                           L.ptr(residual), residual.stride(-2) if residual is not None else 0, act, L.ptr(out),
                                           out.stride(0), L.stream()), 'gemm_scaled')
    return out


def gsf_workspace_floats(clips, clip_len, h, w, fold):
    return int(L.load().tdeed_gsf_workspace_floats(clips, clip_len, h, w, fold))


def gsf_interleaved_positions(fold):
    """position[ch] of input channel ch in the reference's interleaved output (model/impl/gsf.py:84-92)."""
    lib = L.load()
    return [int(lib.tdeed_gsf_interleaved_position(fold, ch)) for ch in range(fold)]


def gsf(x, clips, clip_len, fold, mode, p, workspace, out, natural=False):
    """x NHWC (clips*clip_len, h, w, c) -> out (N*h*w, ld_out) holding the gate-shifted fold channels (natural=True: in input
    channel order, the interleave left to the weight columns of the following 1x1 conv)."""
    n, h, w, c = x.shape
    fn = L.load().tdeed_gsf_fwd_natural if natural else L.load().tdeed_gsf_fwd
    L.check(fn(L.dtype_code(x.dtype), mode, L.ptr(x), clips, clip_len, h, w, c, fold,
               L.ptr(p['bn_scale']), L.ptr(p['bn_shift']), L.ptr(p['w3d']), L.ptr(p['b3d']),
               L.ptr(p.get('cc_w')), L.ptr(p.get('cc_b')), L.ptr(workspace), L.ptr(out),
               out.shape[-1], L.stream()), 'gsf')
    return out


def pool_posenc(x, clip_len, temp_enc, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    L.check(L.load().tdeed_pool_posenc_fwd(L.dtype_code(x.dtype), L.ptr(x), n, h * w, c, clip_len, L.ptr(temp_enc),
                                           L.ptr(out), L.stream()), 'pool_posenc')
    return out


def _fill(struct, tensors):
    for name in struct._names:
        setattr(struct, name, L.ptr(tensors[name]))
    return struct


def sgp_mix(x, t_out, ks, up, w, g_dtype):
    b, t_in, c = x.shape
    y = torch.empty((b, t_out, c), dtype=torch.float32, device=x.device)
    g = torch.empty((b, t_out, c), dtype=g_dtype, device=x.device)
    ws = _fill(L.SgpWeights(), w)
    scratch = torch.empty(int(L.load().tdeed_sgp_mix_workspace_floats(b, t_out, c)), dtype=torch.float32, device=x.device)
    L.check(L.load().tdeed_sgp_mix_fwd(L.ptr(x), b, t_in, t_out, c, ks, up, ctypes.byref(ws), L.ptr(scratch), L.ptr(y), L.ptr(g),
                                       L.dtype_code(g_dtype), L.stream()), 'sgp_mix')
    return y, g


def sgp_mixer_mix(x_coarse, skip, ks, up, w, cat_dtype):
    b, tc, c = x_coarse.shape
    t = skip.shape[1]
    cat = torch.empty((b * t, 6 * c), dtype=cat_dtype, device=skip.device)
    ws = _fill(L.MixerWeights(), w)
    scratch = torch.empty(int(L.load().tdeed_sgp_mixer_workspace_floats(b, tc, t, c)), dtype=torch.float32, device=skip.device)
    L.check(L.load().tdeed_sgp_mixer_mix_fwd(L.ptr(x_coarse), L.ptr(skip), b, tc, t, c, ks, up, ctypes.byref(ws), L.ptr(scratch),
                                             L.ptr(cat), L.dtype_code(cat_dtype), L.stream()), 'sgp_mixer_mix')
    return cat


def groupnorm(x, gamma, beta, out_dtype, groups=16):
    b, t, c = x.shape
    out = torch.empty((b, t, c), dtype=out_dtype, device=x.device)
    scratch = torch.empty(int(L.load().tdeed_groupnorm_workspace_floats(b, t, c, groups)), dtype=torch.float32, device=x.device)
    L.check(L.load().tdeed_groupnorm_fwd(L.ptr(x), b, t, c, groups, L.ptr(gamma), L.ptr(beta), L.ptr(scratch), L.ptr(out),
                                         L.dtype_code(out_dtype), L.stream()), 'groupnorm')
    return out


def heads(feat, w_cls, b_cls, w_displ, b_displ, k_softmax):
    b, t, c = feat.shape
    k_out = w_cls.shape[0]
    logits = torch.empty((b, t, k_out), dtype=torch.float32, device=feat.device)
    probs = torch.empty((b, t, k_softmax), dtype=torch.float32, device=feat.device)
    displ = torch.empty((b, t), dtype=torch.float32, device=feat.device) if w_displ is not None else None
    L.check(L.load().tdeed_heads_fwd(L.ptr(feat), b, t, c, L.ptr(w_cls), L.ptr(b_cls), k_out, L.ptr(w_displ),
                                     L.ptr(b_displ), k_softmax, L.ptr(logits), L.ptr(displ), L.ptr(probs),
                                     L.stream()), 'heads')
    return logits, displ, probs


def softmax_scatter(logits, displ, k_softmax):
    """process_prediction / process_double_head: logits (B,T,K') fp32, displ (B,T) fp32 | None -> probs (B,T,k)."""
    b, t, kk = logits.shape
    probs = torch.empty((b, t, k_softmax), dtype=torch.float32, device=logits.device)
    L.check(L.load().tdeed_softmax_scatter_fwd(L.ptr(logits), kk, L.ptr(displ), b, t, k_softmax, L.ptr(probs),
                                               L.stream()), 'softmax_scatter')
    return probs


def clip_accumulate(scores, support, pred, starts, mode):
    """starts: int32 device tensor, or a host sequence of ints (<= MAX_STARTS_PER_CALL: passed by value, no upload)."""
    video_len, k = scores.shape
    n_clips, t, _ = pred.shape
    if isinstance(starts, torch.Tensor):
        L.check(L.load().tdeed_clip_accumulate(L.ptr(scores), L.ptr(support), video_len, k, L.ptr(pred), L.ptr(starts),
                                               n_clips, t, mode, L.stream()), 'clip_accumulate')
        return
    starts = [int(s) for s in starts]
    assert len(starts) == n_clips
    for lo in range(0, n_clips, L.MAX_STARTS_PER_CALL):            # chunks run in order: same accumulation order
        part = starts[lo:lo + L.MAX_STARTS_PER_CALL]
        arr = (ctypes.c_int * len(part))(*part)
        L.check(L.load().tdeed_clip_accumulate_host(L.ptr(scores), L.ptr(support), video_len, k, L.ptr(pred[lo:lo + len(part)]), arr,
                                                    len(part), t, mode, L.stream()), 'clip_accumulate_host')


def extract_events(scores, support, threshold):
    """Normalises scores/support in place.  Returns dict of device tensors (capacity-sized) + counts."""
    video_len, k = scores.shape
    dev = scores.device
    i32 = dict(dtype=torch.int32, device=dev)
    out = {
        'pred': torch.empty(video_len, **i32),
        'ev_frame': torch.empty(video_len, **i32), 'ev_label': torch.empty(video_len, **i32),
        'ev_score': torch.empty(video_len, dtype=torch.float32, device=dev),
        'hr_frame': torch.empty(video_len * (k - 1), **i32), 'hr_label': torch.empty(video_len * (k - 1), **i32),
        'hr_score': torch.empty(video_len * (k - 1), dtype=torch.float32, device=dev),
        'counts': torch.zeros(2, **i32),
    }
    L.check(L.load().tdeed_extract_events(L.ptr(scores), L.ptr(support), video_len, k, float(threshold),
                                          L.ptr(out['pred']), L.ptr(out['ev_frame']), L.ptr(out['ev_label']),
                                          L.ptr(out['ev_score']), L.ptr(out['hr_frame']), L.ptr(out['hr_label']),
                                          L.ptr(out['hr_score']), L.ptr(out['counts']), L.stream()), 'extract_events')
    return out


def nms(frame, label, score, n_events_dev, k, window, threshold, soft):
    """(soft-)NMS of one video's events.  Returns (out_frame i32, out_label i32, out_score f64, out_count i32[1])."""
    cap = frame.numel()
    dev = frame.device
    ws = torch.empty(int(L.load().tdeed_nms_workspace_bytes(cap, k)), dtype=torch.uint8, device=dev)
    of = torch.empty(cap, dtype=torch.int32, device=dev)
    ol = torch.empty(cap, dtype=torch.int32, device=dev)
    os_ = torch.empty(cap, dtype=torch.float64, device=dev)
    oc = torch.zeros(1, dtype=torch.int32, device=dev)
    L.check(L.load().tdeed_nms(L.ptr(frame), L.ptr(label), L.ptr(score), L.ptr(n_events_dev), cap, k, int(window),
                               float(threshold), int(bool(soft)), L.ptr(ws), L.ptr(of), L.ptr(ol), L.ptr(os_),
                               L.ptr(oc), L.stream()), 'nms')
    return of, ol, os_, oc


def gather_rows(src, src_idx, dst, dst_idx=None, pad_row=None):
    """dst[dst_idx[i] | i] = src[src_idx[i]] (pad_row where src_idx[i] < 0): frame-feature cache plumbing.
    src / dst: contiguous tensors whose leading dim indexes rows; src_idx / dst_idx: int32 device tensors."""
    n = src_idx.numel()
    row_bytes = src[0].numel() * src.element_size()
    assert dst[0].numel() * dst.element_size() == row_bytes and src.is_contiguous() and dst.is_contiguous()
    L.check(L.load().tdeed_gather_rows(L.ptr(src), L.ptr(pad_row), L.ptr(dst), L.ptr(src_idx), L.ptr(dst_idx), n, row_bytes,
                                       L.stream()), 'gather_rows')
    return dst


def match_events(pred_frame, pred_off, gt_frame, gt_off, tolerances, total_pred=None):
    """Greedy matching of util/score.py:45-89 for all (class, video) units x tolerances; int32 device tensors in,
    tp uint8 (n_tol, total_pred) out (see include/tdeed_b200.h (13))."""
    dev = pred_off.device
    n_units = pred_off.numel() - 1
    total_pred = pred_frame.numel() if total_pred is None else total_pred
    total_gt = gt_frame.numel()
    n_tol = tolerances.numel()
    tp = torch.zeros((n_tol, total_pred), dtype=torch.uint8, device=dev)
    if n_units > 0:
        ws = torch.empty((n_tol, max(total_gt, 1)), dtype=torch.uint8, device=dev)
        L.check(L.load().tdeed_match_events(L.ptr(pred_frame), L.ptr(pred_off), L.ptr(gt_frame), L.ptr(gt_off), n_units, total_pred,
                                            total_gt, L.ptr(tolerances), n_tol, L.ptr(ws), L.ptr(tp), L.stream()), 'match_events')
    return tp


def scatter_rows_ring(src, n_rows, ring, first_slot):
    """ring[(first_slot + i) % len(ring)] = src[i] for i < n_rows (frame-feature ring of the video engine)."""
    row_bytes = src[0].numel() * src.element_size()
    assert ring[0].numel() * ring.element_size() == row_bytes and src.is_contiguous() and ring.is_contiguous()
    L.check(L.load().tdeed_scatter_rows_ring(L.ptr(src), L.ptr(ring), n_rows, int(first_slot), ring.shape[0], row_bytes, L.stream()),
            'scatter_rows_ring')


def gather_clip_rows(ring, pad_row, dst, clip_len, first_slots, los, his):
    """dst[b*T + t] = ring[(first_slots[b] + t) % len(ring)] if los[b] <= t < his[b] else pad_row; host int sequences."""
    n = len(first_slots)
    row_bytes = ring[0].numel() * ring.element_size()
    assert dst[0].numel() * dst.element_size() == row_bytes and dst.shape[0] >= n * clip_len and dst.is_contiguous()
    for lo in range(0, n, L.MAX_CLIPS_PER_CALL):
        m = min(L.MAX_CLIPS_PER_CALL, n - lo)
        a = (ctypes.c_int * m)(*[int(v) for v in first_slots[lo:lo + m]])
        b = (ctypes.c_int * m)(*[int(v) for v in los[lo:lo + m]])
        c = (ctypes.c_int * m)(*[int(v) for v in his[lo:lo + m]])
        L.check(L.load().tdeed_gather_clip_rows(L.ptr(ring), L.ptr(pad_row), L.ptr(dst[lo * clip_len:]), m, clip_len, a, b, c,
                                                ring.shape[0], row_bytes, L.stream()), 'gather_clip_rows')
