/* tdeed_b200.h — C-ABI of libtdeed_sm100.so: hand-written sm_100a kernels for T-DEED's per-clip
 * inference/training hot path (BASELINE.json north_star).
 *
 * The reference (arturxe2/T-DEED) is pure Python/PyTorch and has NO operator/plugin/FFI layer
 * (SURVEY.md §8b): its seam is the Python class API of model/model.py.  This header is therefore the
 * boundary a maintainer would bind from the reference's own modules (ctypes stub in INTEGRATION.md);
 * each entry point cites the reference lines whose arithmetic it replaces (paths under the
 * reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory
 *     (the library never allocates or frees device memory and keeps no mutable global state);
 *   - calls are asynchronous on `stream` (a cudaStream_t), re-entrant, and never synchronise;
 *   - activations are channels-last: images NHWC [frames, H, W, C], sequences [B, T, C];
 *   - dtype arguments: TDEED_F32 or TDEED_BF16 (storage type; accumulation is always fp32);
 *   - return 0 on success or a negative tdeed_status; tdeed_last_error() gives the text (thread-local);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TDEED_B200_H_
#define TDEED_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TDEED_OK = 0,
  TDEED_ERR_SHAPE = -1,       /* bad shape / alignment / null pointer */
  TDEED_ERR_UNSUPPORTED = -2, /* configuration outside what the kernels cover */
  TDEED_ERR_CUDA = -3         /* launch or driver error (cudaGetLastError text preserved) */
} tdeed_status;

enum { TDEED_F32 = 0, TDEED_BF16 = 1, TDEED_U8 = 2 };
enum { TDEED_ACT_NONE = 0, TDEED_ACT_RELU = 1, TDEED_ACT_GELU = 2 };
enum { TDEED_SHIFT_GSM = 0, TDEED_SHIFT_GSF = 1 };
enum { TDEED_GEMM_AUTO = 0, TDEED_GEMM_SIMT = 1, TDEED_GEMM_TCGEN05 = 2, TDEED_GEMM_TCGEN05_THIN = 3 };

int tdeed_abi_version(void);
const char* tdeed_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * (1) preprocessing fused into the stem.  Replaces model/model.py:107 (x/255), :121-129 (center /
 * fixed crop, optional horizontal flip, per-clip ImageNet Normalize loop) and timm's
 * stem ConvNormAct(3->32, k3, s2, pad1)+BN(eval)+ReLU.
 * frames: planar [n_frames, 3, in_h, in_w] of u8 or f32 valued 0..255.  The crop window is
 * [crop_y, crop_y+h) x [crop_x, crop_x+w); flip mirrors the cropped window.  w/b are the conv
 * weights [32][3][3][3] and bias [32] with BatchNorm folded in (fp32).
 * out: NHWC [n_frames, ceil(h/2), ceil(w/2), 32] of out_dtype. */
int tdeed_stem_fwd(const void* frames, int frames_dtype, int n_frames, int in_h, int in_w,
                   int crop_y, int crop_x, int h, int w, int flip,
                   const float* weight, const float* bias,
                   void* out, int out_dtype, void* stream);

/* (1b) bf16 tensor-core variant of (1), optionally fused with the first 1x1 conv of stage 1 (timm
 * s1.b1.conv1: 32 -> n1 channels, BN folded, ReLU): the stem is an implicit GEMM (K = 27 padded to 32) whose A
 * operand is built by the threads in shared memory and multiplied by tcgen05.mma; with w1 != NULL the ReLU'd
 * stem rows feed a second MMA and the full-resolution stem activation is never written to HBM.
 * w0: bf16 [32][32] (k = ci*9 + ky*3 + kx, zero padded), b0 fp32 [32]; w1: bf16 [ceil16(n1)][32] zero padded
 * rows, b1 fp32 [n1] (n1 % 8 == 0, n1 <= 64).  out_stem (may be NULL): bf16 NHWC
 * [n, ceil(oh/stem_sub), ceil(ow/stem_sub), 32] holding every stem_sub-th pixel (stem_sub = 2 is what the
 * stride-2 shortcut conv of s1.b1 reads); out_c1: bf16 NHWC [n, oh, ow, n1] (required iff w1). */
int tdeed_stem_tc_fwd(const void* frames, int frames_dtype, int n_frames, int in_h, int in_w,
                      int crop_y, int crop_x, int h, int w, int flip,
                      const void* w0_bf16, const float* b0, const void* w1_bf16, const float* b1, int n1,
                      void* out_stem, int stem_sub, void* out_c1, void* stream);

/* (1c) second-generation fused stem for uint8 frames (stem_tc2.cu): shifted-descriptor implicit GEMM on RAW pixel values — the
 * normalisation is folded into the weights, zero padding becomes padding with the raw value 255*mean (pad_rgb_host, 3 floats),
 * two taps share one K=16 MMA.  wimg: tdeed_stem_tc2_wimg_bytes() bytes = 5 B tiles [32 out][16 k] (canonical K-major
 * no-swizzle; k 0..2 = channels of the pair's first tap, k 8..10 = second tap; pairs (0,0)+(0,2), (2,0)+(2,2), (0,1)+(2,1),
 * (1,0)+(1,2), (1,1)) of  w * bn_scale / (255 * std);  b0 = bn_shift - sum_k bf16(w') * 255 * mean.  w1 / b1 / n1 / out_stem /
 * stem_sub / out_c1 as in (1b); the fused conv1 is mandatory here. */
long long tdeed_stem_tc2_wimg_bytes(void);
int tdeed_stem_tc2_fwd(const void* frames_u8, int n_frames, int in_h, int in_w, int crop_y, int crop_x, int h, int w,
                       int flip, const void* wimg, const float* b0, const float* pad_rgb_host, const void* w1_bf16,
                       const float* b1, int n1, void* out_stem, int stem_sub, void* out_c1, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (2) 1x1 convolution / linear layer as a GEMM with fused epilogue:
 *        out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] + residual[m, n] )
 * Replaces timm Bottleneck conv1 / conv3 / downsample (+folded BN, ReLU, shortcut add), the SGP MLP
 * and concat_fc 1x1 Conv1d layers (model/modules.py:134-138,186,244-251,307-309) and nn.Linear.
 * A is given as up to TDEED_GEMM_MAX_SEGS column segments that are concatenated along K (the
 * GatedShift "y[:, :fold] = gs(x[:, :fold]); y[:, fold:] = x[:, fold:]" of model/shift.py:89-93 is a
 * virtual concat of the gate-shift output and the untouched channels).  Segment s covers columns
 * [col0, col0+k) of the row-major matrix `a` with leading dimension `lda` (elements); segments
 * consume W's columns in order.  When gather_stride > 1, segment rows are a strided spatial
 * subsample of an NHWC tensor [frames, gather_h, gather_w, lda]: row m = (f, oy, ox) reads pixel
 * (f, oy*stride, ox*stride) — the stride-2 1x1 "downsample" shortcut.
 * W: [N, K] row-major (K = sum of segment k) of `dtype`.  bias: fp32 [N] or NULL.  residual: [M, ldr]
 * of res_dtype or NULL.  out: [M, ldo] of out_dtype.  lda/ldr/ldo must be multiples of 8 elements
 * and all base pointers 16-byte aligned.  backend: TDEED_GEMM_AUTO picks, for bf16, the thin-K tcgen05 kernel
 * (cp.async producers, resident weights) when K <= 64 and N <= 256 and the TMA-fed persistent tcgen05 kernel otherwise;
 * for f32 the exact CUDA-core kernel. */
#define TDEED_GEMM_MAX_SEGS 2
typedef struct {
  const void* a;
  long long lda;
  int col0;
  int k;
} tdeed_gemm_seg;

int tdeed_gemm_fwd(int dtype, long long M, int N, int nseg, const tdeed_gemm_seg* segs_host,
                   int gather_stride, int gather_h, int gather_w,
                   const void* W, const float* bias,
                   const void* residual, long long ldr, int res_dtype,
                   int act, void* out, long long ldo, int out_dtype, int backend, void* stream);

/* (2c) conv3 of a RegNetY bottleneck with the squeeze-excite gate folded into its A operand (timm Bottleneck.forward:
 * x = conv3(se(conv2(...)))):  out[m, n] = act( sum_k (A[m, k] * a_scale[m / scale_rows][k]) * W[n, k] + bias[n] + residual[m, n] ).
 * A, W, residual, out bf16; a_scale fp32 [M / scale_rows][K] (the gate of tdeed_se_gate_fwd; scale_rows = pixels per frame).
 * The product A * gate is rounded to bf16 exactly like the stand-alone scale pass of tdeed_se_fwd, inside shared memory between
 * the TMA load and the tcgen05 MMA.  bf16 tcgen05 backend only (N tile <= 256 per accumulator; act none | relu). */
int tdeed_gemm_scaled_fwd(long long M, int N, int K, const void* A, long long lda, const float* a_scale, int scale_rows,
                          const void* W, const float* bias, const void* residual, long long ldr, int act, void* out,
                          long long ldo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (3) grouped 3x3 convolution (+folded BN + ReLU): timm Bottleneck conv2, group width 8 / 16,
 * stride 1 or 2, pad 1.  in: NHWC [n, h, w, c]; weight fp32 [c][group_width][3][3] (torch layout,
 * BN folded), bias fp32 [c]; out: NHWC [n, ceil(h/stride), ceil(w/stride), c]. */
int tdeed_conv3x3g_fwd(int dtype, const void* in, int n, int h, int w, int c, int group_width, int stride,
                       const float* weight, const float* bias, void* out, void* stream);

/* (3b) tcgen05 variant of (3) for bf16 activations.  wimg: the weights as UMMA B tiles, bf16
 * [ceil(c/16)][9 taps][16 out][16 in] per 16-channel pair (block-diagonal for group width 8), each tile stored in
 * the canonical K-major no-swizzle layout [n/8][k/8][n%8][k%8] (512 B) — see tdeed_b200/engine.py:conv3_weight_image. */
int tdeed_conv3x3g_tc_fwd(const void* in, int n, int h, int w, int c, int stride, const void* wimg,
                          const float* bias, void* out, void* stream);

/* (4) squeeze-excite, in place: x *= sigmoid(fc2(relu(fc1(mean_hw(x))))).  timm SEModule.
 * x: NHWC [n, hw, c]; w1 [rd][c], b1 [rd], w2t [rd][c] (fc2 weight TRANSPOSED so that channel reads coalesce),
 * b2 [c]; all fp32.  workspace: fp32, tdeed_se_workspace_floats(n, c) elements (per-frame means and scales). */
long long tdeed_se_workspace_floats(int n, int c);
int tdeed_se_fwd(int dtype, void* x, int n, int hw, int c, int rd,
                 const float* w1, const float* b1, const float* w2t, const float* b2, float* workspace, void* stream);
/* gate only (mean -> fc1 -> ReLU -> fc2 -> sigmoid): the gate [n, c] fp32 is left at workspace + n*c for a consumer that applies
 * it itself (tdeed_gemm_scaled_fwd); x is not modified. */
int tdeed_se_gate_fwd(int dtype, const void* x, int n, int hw, int c, int rd, const float* w1, const float* b1,
                      const float* w2_t, const float* b2, float* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (5) Gate-Shift module on the first `fold` channels of x (model/shift.py:64-93):
 *   GSM  model/impl/gsm.py:89-116, GSF model/impl/gsf.py:38-93  (eval-mode BatchNorm3d folded into
 *   bn_scale/bn_shift).  x: NHWC [clips*clip_len, h, w, c]; out: [clips*clip_len*h*w, ld_out] holding the
 *   gate-shifted `fold` channels (channel-interleaved as the reference), ld_out >= fold, multiple of 8.
 *   conv3d_w fp32 [2][fold/2][3][3][3], conv3d_b [2]; cc_w fp32 [2][2][3][3] = channel_conv1/2 weights,
 *   cc_b [2] (ignored for GSM).  workspace: fp32, at least tdeed_gsf_workspace_floats(...) elements. */
long long tdeed_gsf_workspace_floats(int clips, int clip_len, int h, int w, int fold);
int tdeed_gsf_fwd(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                  const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                  const float* cc_w, const float* cc_b, float* workspace,
                  void* out, int ld_out, void* stream);
/* Same, but out[:, ch] holds the result for INPUT channel ch (no interleave).  The interleave is a fixed channel permutation in
 * front of the block's 1x1 convolution, so a caller that owns that convolution's weight permutes its columns once instead:
 * out_natural[:, ch] == out_interleaved[:, tdeed_gsf_interleaved_position(fold, ch)].  (The bf16 kernel then stores 16 bytes
 * per thread straight from registers.) */
int tdeed_gsf_fwd_natural(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                          const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                          const float* cc_w, const float* cc_b, float* workspace,
                          void* out, int ld_out, void* stream);
int tdeed_gsf_interleaved_position(int fold, int ch);

/* (6) global average pool + positional encoding: feat[f, :] = mean_hw(x[f]) + temp_enc[f % clip_len, :]
 * (timm head global_pool, model/model.py:133-137).  x: NHWC [n, hw, c]; temp_enc fp32 [clip_len, c];
 * out fp32 [n, c]. */
int tdeed_pool_posenc_fwd(int dtype, const void* x, int n, int hw, int c, int clip_len,
                          const float* temp_enc, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (7) SGP token mixing (model/modules.py:159-186, everything of SGPBlock.forward before the MLP), on
 * [B, T, C] fp32 sequences.  If t_in != t_out the input is first AdaptiveMaxPool1d'ed to t_out
 * (model/modules.py:64,76).  With xp = (pooled) x and ln = LayerNorm_C(xp):
 *    y = xp + fc(ln)*relu(global_fc(mean_T ln)) + (convw(ln)+convkw(ln))*psi(ln) + ln
 *    g = GroupNorm16(y)
 * Depthwise weights fp32: psi_w/convw_w [C][ks], convkw_w [C][up], fc_w/gfc_w [C]; biases [C].
 * y: fp32 [B, t_out, C];  g: [B, t_out, C] of g_dtype (the MLP's GEMM operand). */
typedef struct {
  const float *ln_w, *ln_b, *gn_w, *gn_b;
  const float *psi_w, *psi_b, *fc_w, *fc_b, *convw_w, *convw_b, *convkw_w, *convkw_b, *gfc_w, *gfc_b;
} tdeed_sgp_weights;
long long tdeed_sgp_mix_workspace_floats(int B, int t_out, int C);
int tdeed_sgp_mix_fwd(const float* x, int B, int t_in, int t_out, int C, int ks, int up,
                      const tdeed_sgp_weights* w_host, float* workspace, float* y, void* g, int g_dtype, void* stream);
/* workspace: >= tdeed_sgp_mix_workspace_floats(B, t_out, C) floats, 16-byte aligned (LayerNorm row statistics, partial
 * column sums for the phi gate, partial GroupNorm sums in double: the clip-wide reductions are two-level and deterministic). */

/* (8) SGPMixer token mixing (model/modules.py:283-307): z = LN1(skip) [B,T,C], xu =
 * linear-upsample(align_corners) of LN2(x) [B,t_coarse,C] to T; writes the 6C-wide concat
 * [out1,out2,out3,out4,z,xu] as [B*T, 6C] of cat_dtype for the concat_fc GEMM. */
typedef struct {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const float *psi1_w, *psi1_b, *psi2_w, *psi2_b, *convw1_w, *convw1_b, *convkw1_w, *convkw1_b;
  const float *convw2_w, *convw2_b, *convkw2_w, *convkw2_b;
  const float *fc1_w, *fc1_b, *gfc1_w, *gfc1_b, *fc2_w, *fc2_b, *gfc2_w, *gfc2_b;
} tdeed_mixer_weights;
long long tdeed_sgp_mixer_workspace_floats(int B, int t_coarse, int T, int C);
int tdeed_sgp_mixer_mix_fwd(const float* x_coarse, const float* skip, int B, int t_coarse, int T, int C,
                            int ks, int up, const tdeed_mixer_weights* w_host, float* workspace,
                            void* cat, int cat_dtype, void* stream);

/* (9) GroupNorm(16, C) over [B, T, C] fp32 -> out of out_dtype (model/modules.py:115,186,311). */
long long tdeed_groupnorm_workspace_floats(int B, int T, int C, int groups);
int tdeed_groupnorm_fwd(const float* x, int B, int T, int C, int groups, const float* gamma, const float* beta,
                        float* workspace, void* out, int out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (10) heads + softmax + displacement scatter-max.  Replaces FCLayers/FC2Layers (eval: dropout is
 * identity; model/modules.py:366-387, model/model.py:141-146), torch.softmax and the B x T Python
 * loop of process_prediction / process_double_head (model/modules.py:406-426).
 * feat fp32 [B,T,C]; w_cls fp32 [k_out, C], b_cls [k_out]; w_displ fp32 [C] / b_displ [1] or NULL.
 * logits fp32 [B,T,k_out]; displ fp32 [B,T] (if w_displ); probs fp32 [B,T,k_softmax]:
 *   p = softmax(logits[..., :k_softmax]);  probs[b, clamp(t - rint(displ[b,t]), 0, T-1)] = max(., p[b,t])
 *   (plain softmax when w_displ is NULL). */
int tdeed_heads_fwd(const float* feat, int B, int T, int C, const float* w_cls, const float* b_cls, int k_out,
                    const float* w_displ, const float* b_displ, int k_softmax,
                    float* logits, float* displ, float* probs, void* stream);

/* (10b) process_prediction / process_double_head on precomputed logits (model/modules.py:406-426):
 * probs[b, clamp(t - rint(displ[b,t]), 0, T-1)] = max(., softmax(logits[b,t,:k_softmax]));  displ may be
 * NULL (plain softmax).  logits fp32 [B,T,ld_logits]; probs fp32 [B,T,k_softmax]. */
int tdeed_softmax_scatter_fwd(const float* logits, int ld_logits, const float* displ, int B, int T,
                              int k_softmax, float* probs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (11) per-video post-processing, bit-exact with util/eval.py.
 * clip_accumulate (util/eval.py:303-349): for clips i = 0..n_clips-1 IN ORDER,
 *   scores[start_i + t] += pred[i, t];  support += (rowsum != 0) [mode 0, batched path]  or += 1 [mode 1,
 *   TTA path]; rows before frame 0 / past video_len are dropped.  pred fp32 [n_clips, T, K]; starts i32. */
int tdeed_clip_accumulate(float* scores, int* support, int video_len, int K,
                          const float* pred, const int* starts, int n_clips, int T, int mode, void* stream);
/* same with the clip starts given as a HOST array (n_clips <= TDEED_MAX_STARTS_PER_CALL; passed to the kernel by value) */
#define TDEED_MAX_STARTS_PER_CALL 256
int tdeed_clip_accumulate_host(float* scores, int* support, int video_len, int K,
                               const float* pred, const int* starts_host, int n_clips, int T, int mode, void* stream);

/* extract (util/eval.py:87-193): support==0 -> 1; scores /= support; pred = argmax; events = frames with
 * pred != 0; high-recall events = every (frame, class>=1) with score >= threshold (fp32 compare),
 * frame-major then class order.  counts_out i32 [2] = {n_events, n_high_recall}.  Capacities:
 * ev_* video_len entries, hr_* video_len*(K-1) entries. */
int tdeed_extract_events(float* scores, int* support, int video_len, int K, float threshold,
                         int* pred, int* ev_frame, int* ev_label, float* ev_score,
                         int* hr_frame, int* hr_label, float* hr_score, int* counts_out, void* stream);

/* (soft-)NMS of one video's event list (util/eval.py:195-261).  Input events must be in the
 * reference's frame-major order with at most one event per (frame, label); labels in [1, K).
 * n_events_dev points to the device-side event count (e.g. counts_out+1 of tdeed_extract_events), so
 * no host round trip is needed; capacity bounds the arrays.  soft=0: hard NMS, out_score holds the
 * fp32 scores widened to f64;  soft=1: quadratic-decay soft NMS in f64 exactly as the reference.
 * threshold is compared in double (util/eval.py:213,247).  Output is sorted by frame, ties in the
 * reference's label-bucket (first-appearance) order; out_count i32 [1].
 * workspace: bytes >= tdeed_nms_workspace_bytes(capacity, K). */
long long tdeed_nms_workspace_bytes(int capacity, int K);
int tdeed_nms(const int* frame, const int* label, const float* score, const int* n_events_dev, int capacity,
              int K, int window, double threshold, int soft, void* workspace,
              int* out_frame, int* out_label, double* out_score, int* out_count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (12) frame-feature cache of the video-level engine.  The reference's inference clips overlap by 75 %
 * (dataset/frame.py:409-423: starts every (clip_len - overlap_len) * stride frames; util/eval.py:289-349 runs
 * each clip through the whole network), while everything before the first GatedShift (model/shift.py:47-59: s3.b1)
 * is clip-independent.  Those per-frame features are computed once per unique frame and kept in HBM; this entry
 * point files rows into the cache and assembles clip batches from it:
 *     dst[dst_idx ? dst_idx[i] : i] = (src_idx[i] < 0) ? pad_row : src[src_idx[i]]      for i in [0, n_rows)
 * rows are row_bytes (multiple of 16) long and 16-byte aligned.  Negative source indices select pad_row — the
 * features of the all-zero frame the reference pads clips with before frame 0 / past the end of the video
 * (dataset/frame.py:622-625); pad_row may be NULL when no index is negative. */
int tdeed_gather_rows(const void* src, const void* pad_row, void* dst, const int* src_idx, const int* dst_idx,
                      int n_rows, long long row_bytes, void* stream);
/* (12b) the two steady-state maps of the video engine without index tensors (the integers travel as kernel parameters,
 * so no small host->device copy competes with the frame uploads for the DMA engine):
 *   scatter_rows_ring:  ring[(first_slot + i) % ring_slots] = src[i]                                   i in [0, n_rows)
 *   gather_clip_rows:   dst[b*T + t] = (lo[b] <= t < hi[b]) ? ring[(first_slot[b] + t) % ring_slots] : pad_row
 *                       — clip b of T frames whose frame 0 sits in ring slot first_slot[b]; frames outside [lo, hi) lie before
 *                       frame 0 / past the end of the video (dataset/frame.py:412-423,622-625: zero padding).
 * *_host arrays are HOST pointers with n_clips <= TDEED_MAX_CLIPS_PER_CALL entries. */
#define TDEED_MAX_CLIPS_PER_CALL 128
int tdeed_scatter_rows_ring(const void* src, void* ring, int n_rows, int first_slot, int ring_slots, long long row_bytes,
                            void* stream);
int tdeed_gather_clip_rows(const void* ring, const void* pad_row, void* dst, int n_clips, int T, const int* first_slot_host,
                           const int* lo_host, const int* hi_host, int ring_slots, long long row_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (13) greedy prediction <-> ground-truth matching of the mAP scorer (util/score.py:45-89,
 * compute_average_precision).  A unit = the predictions of one (class, video) in descending-score order
 * (pred_frame[pred_off[u] .. pred_off[u+1])) and that video's ground-truth frames of the class in label-file order
 * (gt_frame[gt_off[u] .. gt_off[u+1])).  For every tolerance independently, predictions are visited in order; each takes
 * the closest ground-truth frame not recalled yet (lowest list index among equal distances) and, if it lies within the
 * tolerance, recalls it (all duplicates of that frame value with it): tp[t][p] = 1, else 0.  tp is u8 [n_tol][total_pred]
 * (entries of predictions that belong to no unit are left untouched: zero them beforehand);
 * recalled_ws: u8 [n_tol][total_gt] scratch. */
int tdeed_match_events(const int* pred_frame, const int* pred_off, const int* gt_frame, const int* gt_off, int n_units,
                       int total_pred, int total_gt, const int* tolerances, int n_tol, unsigned char* recalled_ws,
                       unsigned char* tp, void* stream);

#ifdef __cplusplus
}
#endif

#include "tdeed_b200_train.h" /* training-step entry points */

#endif /* TDEED_B200_H_ */
