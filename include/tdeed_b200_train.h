/* tdeed_b200_train.h — training-step entry points of libtdeed_sm100.so (included by tdeed_b200.h; same conventions:
 * device pointers, caller-owned memory and workspaces, asynchronous on `stream`, 0 / negative tdeed_status).
 *
 * The reference trains through PyTorch autograd (model/model.py:193-332, model/modules.py:388-401): there is no
 * operator boundary to mirror, so every function names the reference lines whose forward (in train mode) or whose
 * autograd-derived backward it computes.  Gradients are WRITTEN (not accumulated) unless an `add` operand is named;
 * all reductions use a fixed order (bitwise deterministic, no atomics).
 */
#ifndef TDEED_B200_TRAIN_H_
#define TDEED_B200_TRAIN_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- training-mode BatchNorm around the convolutions (timm BatchNormAct2d / nn.BatchNorm3d in .train()) ----------
 * x: [M, ld] rows of C channels (ld >= C; C may be a column slice).  stats: fp32 [4][C] = mean, invstd, scale =
 * gamma*invstd, shift = beta - mean*scale.  running_mean/var (nullable pair) get the momentum update with the unbiased
 * variance.  workspace: tdeed_bn_workspace_floats(C) floats. */
long long tdeed_bn_workspace_floats(int C);
int tdeed_bn_stats(int dtype, const void* x, long long M, int C, long long ld, const float* gamma, const float* beta,
                   float eps, float momentum, float* running_mean, float* running_var, float* stats, float* workspace,
                   void* stream);
/* z = act(y*scale + shift (+ residual)); y, residual, out: [M, C] of dtype. */
int tdeed_bn_act_fwd(int dtype, const void* y, long long M, int C, const float* stats, const void* residual, int relu,
                     void* out, void* stream);
/* g = dz * (z > 0) (z == NULL: no ReLU);  dgamma = sum g*xhat, dbeta = sum g;  dy = scale*(g - dbeta/M - xhat*dgamma/M);
 * dres (nullable) = g, the gradient of the residual operand.  dy may alias dz.  Passing z == y (the same pointer) asks for the
 * ReLU mask to be recomputed as (y*scale + shift > 0) — valid when the forward had no residual — which saves reading z. */
int tdeed_bn_act_bwd(int dtype, const void* dz, const void* z, const void* y, long long M, int C, const float* stats,
                     float* dgamma, float* dbeta, void* dy, void* dres, float* workspace, void* stream);

/* ---- weight-gradient GEMM: out[i, j] = alpha * sum_r A[r, i] * B[r, j]  (A [R, m], B [R, n] row-major) --------------
 * dW = dY^T X of every 1x1 conv / linear layer.  gather_stride > 1: B's row r = (f, oy, ox) reads pixel
 * (f, oy*s, ox*s) of an NHWC tensor [frames, gather_h, gather_w, ldb] (stride-2 shortcut conv).  out fp32 [m, ldo]. */
long long tdeed_gemm_tn_workspace_floats(long long R, int m, int n);
int tdeed_gemm_tn(int a_dtype, const void* A, long long lda, int b_dtype, const void* B, long long ldb, long long R,
                  int m, int n, int gather_stride, int gather_h, int gather_w, float alpha, float* out, long long ldo,
                  float* workspace, void* stream);
long long tdeed_colsum_workspace_floats(long long M, int C);
int tdeed_colsum(int dtype, const void* x, long long M, int C, long long ld, float* out, float* workspace, void* stream);
/* dst[f, oy, ox, :] = src[f, s*oy, s*ox, :]: compact copy of the pixels a stride-s 1x1 conv reads (its dW then runs on the dense
 * tcgen05 GEMM) */
int tdeed_strided_gather(int dtype, const void* src, void* dst, int n, int h, int w, int c, int stride, void* stream);
/* dst[f, s*oy, s*ox, :] += src[f, oy, ox, :]: data gradient of a stride-s 1x1 conv added into the block-input gradient */
int tdeed_strided_add(int dtype, void* dst, const void* src, int n, int h, int w, int c, int stride, void* stream);

/* ---- spatial convolutions ------------------------------------------------------------------------------------------
 * raw (no bias / BN / activation) forwards; weights in the torch layouts of tdeed_b200.h (3). */
int tdeed_stem_raw_fwd(const void* frames, int frames_dtype, int unit_input, int n_frames, int in_h, int in_w,
                       int crop_y, int crop_x, int h, int w, int flip, const float* weight, void* out, int out_dtype,
                       void* stream);
long long tdeed_stem_bwd_weight_workspace_floats(void);
int tdeed_stem_bwd_weight(const void* frames, int frames_dtype, int unit_input, int n_frames, int in_h, int in_w,
                          int crop_y, int crop_x, int h, int w, int flip, const void* dy, int dy_dtype, float* dw,
                          float* workspace, void* stream);
/* im2col of the normalised stem input as bf16 rows [n*oh*ow][32] (k = ci*9 + ky*3 + kx, columns 27..31 zero): the stem weight
 * gradient is then tdeed_gemm_tn(dY [P, 32], patches [P, 32]) on tcgen05 */
int tdeed_stem_im2col(const void* frames, int frames_dtype, int unit_input, int n_frames, int in_h, int in_w, int crop_y,
                      int crop_x, int h, int w, int flip, void* patches_bf16, void* stream);
int tdeed_conv3x3g_raw_fwd(int dtype, const void* in, int n, int h, int w, int c, int group_width, int stride,
                           const float* weight, void* out, void* stream);
int tdeed_conv3x3g_bwd_data(int dtype, const void* dy, int n, int h, int w, int c, int group_width, int stride,
                            const float* weight, void* dx, void* stream);
/* tcgen05 variants for bf16 activations: tdeed_conv3_weight_image turns the fp32 weights into the UMMA B tiles that
 * tdeed_conv3x3g_tc_fwd (tdeed_b200.h 3b) consumes — tdeed_conv3_weight_image_elems(c) bf16 elements; with
 * transpose_flip = 1 the image is that of the transposed, 180-degree-rotated kernel, so that tdeed_conv3x3g_tc_raw_fwd on
 * dy (stride 1) is the data gradient of the stride-1 convolution. */
long long tdeed_conv3_weight_image_elems(int c);
int tdeed_conv3_weight_image(const float* weight, int c, int group_width, int transpose_flip, void* wimg, void* stream);
int tdeed_conv3x3g_tc_raw_fwd(const void* in, int n, int h, int w, int c, int stride, const void* wimg, void* out,
                              void* stream);
/* data gradient of the stride-2 conv on tcgen05: one stride-1 conv over the dy grid per input-pixel parity (py, px), weight
 * images built with tdeed_conv3_weight_image(transpose_flip = 2 + 2*py + px) and stored back to back in wimgs; each scatters into
 * dx [n, h, w, c] at pixels (2j+py, 2i+px). */
int tdeed_conv3x3g_tc_bwd_data_s2(const void* dy, int n, int h, int w, int c, const void* wimgs, void* dx, void* stream);
long long tdeed_conv3x3g_bwd_weight_workspace_floats(int n, int h, int w, int c, int group_width, int stride);
int tdeed_conv3x3g_bwd_weight(int dtype, const void* x, const void* dy, int n, int h, int w, int c, int group_width,
                              int stride, float* dw, float* workspace, void* stream);

/* ---- squeeze-excite (timm SEModule) --------------------------------------------------------------------------------
 * train fwd is out-of-place; its workspace (tdeed_se_workspace_floats) keeps mean [n][c] | scale [n][c] for the backward.
 * bwd writes dx and the per-frame vectors vec = dv [n][c] | dh [n][rd] | h [n][rd] | dm [n][c] | ds [n][c]
 * (tdeed_se_bwd_vec_floats); the fc gradients follow with tdeed_gemm_tn / tdeed_colsum:
 *   d fc2.weight [c][rd] = dv^T h,  d fc2.bias = colsum(dv),  d fc1.weight [rd][c] = dh^T mean,  d fc1.bias = colsum(dh). */
int tdeed_se_train_fwd(int dtype, const void* x, void* out, int n, int hw, int c, int rd, const float* w1, const float* b1,
                       const float* w2t, const float* b2, float* workspace, void* stream);
long long tdeed_se_bwd_vec_floats(int n, int c, int rd);
int tdeed_se_bwd(int dtype, const void* x, const void* du, int n, int hw, int c, int rd, const float* w1, const float* b1,
                 const float* w2t, const float* fwd_workspace, void* dx, float* vec, void* stream);
/* global average pool + temp_enc backward: dz[f, p, :] = dfeat[f, :]/hw;  d temp_enc[t, :] = sum_b dfeat[b*T + t, :] */
int tdeed_pool_posenc_bwd(int dtype, const float* dfeat, int clips, int clip_len, int hw, int c, void* dz,
                          float* d_temp_enc, void* stream);

/* ---- Gate-Shift(-Fuse) with training-mode BatchNorm3d (model/shift.py:64-93, model/impl/gsf.py:38-93) ---------------
 * forward: tdeed_bn_stats on x[:, :fold] then tdeed_gsf_cat_fwd, which writes the full concat
 * [gs(x[:, :fold]) | x[:, fold:]] as [frames*h*w, c]; its workspace (tdeed_gsf_workspace_floats) is kept for the backward.
 * backward: dx [frames*h*w, c] = d/dx of the concat (+ add, nullable, same shape);  d_cc fp32 [2][19] = per group the 18
 * channel_conv weights then its bias;  d_conv3d_w [2][fold/2][27], d_conv3d_b [2], d_bn_gamma/beta [fold]. */
int tdeed_gsf_cat_fwd(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                      const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                      const float* cc_w, const float* cc_b, float* workspace, void* out, void* stream);
long long tdeed_gsf_bwd_workspace_floats(int clips, int clip_len, int h, int w, int fold);
int tdeed_gsf_bwd(int dtype, int mode, const void* x, const void* dcat, const void* add, int clips, int clip_len,
                  int h, int w, int c, int fold, const float* bn_stats, const float* conv3d_w, const float* cc_w,
                  const float* fwd_workspace, float* workspace, void* dx, float* d_conv3d_w, float* d_conv3d_b,
                  float* d_cc, float* d_bn_gamma, float* d_bn_beta, void* stream);

/* ---- ED-SGP-Mixer backward pieces (model/modules.py:58-363), fp32 [B, T, C] ------------------------------------------
 * chan_ln_fwd: (AdaptiveMaxPool1d t_in -> T when t_in > T, then) channel LayerNorm; xp / argmax (nullable) receive the
 * pooled input and the arg-max rows; stats [B*T][2] = mean, rstd. */
int tdeed_chan_ln_fwd(const float* x, int B, int t_in, int T, int C, const float* w, const float* b, float* xp,
                      int* argmax, float* ln, float* stats, void* stream);
int tdeed_chan_ln_bwd(const float* xp, const float* stats, const float* dln, long long ld, int rows, int C,
                      const float* w, const float* add, float* dx, float* dw, float* db, void* stream);
int tdeed_maxpool_bwd(const float* dxp, const int* argmax, int B, int t_in, int T, int C, const float* add, float* dx,
                      void* stream);
/* backward of  fc(ln)*relu(global_fc(mean_T ln)) + (convw(ln) + convkw(ln))*psi(ln) [+ ln]  (SGPBlock :168-184 and each of
 * the two inputs of SGPMixer :291-305).  d_conv / d_fc / d_id: upstream gradients of the three terms with leading dim
 * ld_g (the same pointer three times for SGPBlock; column slices of d(cat) for the mixer; d_id nullable).
 * weights / grads: 10 pointers each, order psi_w, psi_b, convw_w, convw_b, convkw_w, convkw_b, fc_w, fc_b, gfc_w, gfc_b. */
long long tdeed_sgp_branch_bwd_workspace_floats(int B, int T, int C);
int tdeed_sgp_branch_bwd(const float* ln, long long ld_ln, const float* d_conv, const float* d_fc, const float* d_id,
                         long long ld_g, int B, int T, int C, int ks, int up, const float* const* weights_host,
                         float* const* grads_host, float* d_ln, float* workspace, void* stream);
/* dy = add + GroupNorm_backward(dg);  workspace: 2*B*C floats */
int tdeed_groupnorm_bwd(const float* y, const float* dg, int B, int T, int C, int groups, const float* gamma,
                        const float* add, float* dy, float* dgamma, float* dbeta, float* workspace, void* stream);
int tdeed_gelu_fwd(const float* h, long long n, void* out, int out_dtype, void* stream);
int tdeed_gelu_bwd(const float* h, const float* da, long long n, void* dh, int out_dtype, void* stream);
int tdeed_upsample_bwd(const float* dxu, int B, int t_coarse, int T, int C, float* dx, void* stream);
int tdeed_cast_f32(const float* in, long long n, void* out, int out_dtype, void* stream);

/* ---- per-clip training augmentation (model/model.py:77-84,154-157: torchvision ColorJitter / GaussianBlur(5) / hflip) -----------
 * tdeed_aug_color: frames planar [n, 3, in_h, in_w] u8 | f32, cropped to [crop_y, +h) x [crop_x, +w), scaled by in_scale (1/255),
 * then hue shift -> saturation -> brightness (each if *_on), out fp32 planar [n, 3, h, w] in [0, 1].
 * tdeed_aug_gray_mean: mean[n] of the grayscale image (the pivot of adjust_contrast).
 * tdeed_aug_contrast_blur_flip: contrast (pivot mean[frame]) -> 5x5 Gaussian (kernel1d_host: 5 host floats, reflect padding) ->
 * horizontal flip; x and out must not alias. */
/* mixup of two uint8 clip batches (model/model.py:228-254): out[s, i] = fl32(lam[s][0]) * a[s, i] + fl32(lam[s][1]) * b[s, i],
 * the reference's arithmetic (two rounded products, one rounded sum) in one pass; a, b: u8 [n_samples, per_sample] (per_sample a
 * multiple of 16, 16-byte aligned), lam: fp32 [n_samples][2] on the device, out: fp32 [n_samples, per_sample]. */
int tdeed_mixup_u8(const void* a, const void* b, const float* lam, int n_samples, long long per_sample, float* out, void* stream);
int tdeed_aug_color(const void* frames, int frames_dtype, float in_scale, int n_frames, int in_h, int in_w, int crop_y,
                    int crop_x, int h, int w, int hue_on, float hue, int sat_on, float sat, int bri_on, float bri,
                    float* out, void* stream);
int tdeed_aug_gray_mean(const float* x, int n_frames, int hw, float* mean, void* stream);
int tdeed_aug_contrast_blur_flip(const float* x, int n_frames, int h, int w, int con_on, float con, const float* mean,
                                 int blur_on, const float* kernel1d_host, int flip, float* out, void* stream);

/* ---- heads, loss, optimizer ---------------------------------------------------------------------------------------- */
int tdeed_dropout_fwd(const float* x, long long n, float p, unsigned long long seed, float* out, unsigned char* mask,
                      void* stream);
/* same, with the seed read from device memory (seed_dev[0]*2 + salt): a CUDA graph replay then draws a new mask per step */
int tdeed_dropout_fwd_devseed(const float* x, long long n, float p, const long long* seed_dev, unsigned long long salt,
                              float* out, unsigned char* mask, void* stream);
int tdeed_dropout_bwd(const float* dy, const unsigned char* mask, long long n, float p, const float* add, float* dx,
                      void* stream);
int tdeed_linear_fwd(const float* x, int M, int C, const float* W, const float* b, int N, float* out, int ldo, void* stream);
int tdeed_linear_bwd_data(const float* dout, int ldd, int M, int C, const float* W, int N, const float* add, float* dx,
                          void* stream);
/* F.cross_entropy(weight=class_weight) with int64 class targets (weighted mean) or probability targets (mean over rows)
 * + F.mse_loss(displ, labelD).mean()  (model/model.py:308-319).  loss_out fp32 [3] = total, CE, MSE.
 * An int64 target outside [0, K) (its head's range in the two-head variant) makes torch raise; the kernels never index out of
 * bounds (the label is clamped) and report it as a NaN loss (total and CE). */
int tdeed_ce_mse_loss(const float* logits, int M, int K, int ld_logits, const long long* target_hard,
                      const float* target_soft, const float* class_weight, const float* displ, const float* labelD,
                      float* loss_out, float* dlogits, float* ddispl, void* stream);
/* joint-dataset double head (model/model.py:278-306): dataset int32 [B] in {1, 2} selects head 1 (columns [0, n1)) or head 2
 * ([n1, n1+n2)) per sample; per-sample cross entropy (class_weight[:k]) over its T rows, summed / B; int64 targets of dataset-2
 * samples are shifted by n1 as update_labels_2heads leaves them; soft targets are [B*T, n1+n2].  dlogits: [B*T, n1+n2]. */
int tdeed_ce_mse_loss_2heads(const float* logits, int B, int T, int n1, int n2, int ld_logits, const int* dataset,
                             const long long* target_hard, const float* target_soft, const float* class_weight,
                             const float* displ, const float* labelD, float* loss_out, float* dlogits, float* ddispl,
                             void* stream);
/* torch.optim.AdamW single step on a flat fp32 buffer; g is multiplied by grad_scale first; shadow_bf16 (nullable)
 * receives the bf16 copy of the updated weights. */
int tdeed_adamw_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                     double eps, double weight_decay, int step, float grad_scale, void* shadow_bf16, void* stream);
int tdeed_axpy(const float* x, float alpha, long long n, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDEED_B200_TRAIN_H_ */
