// Heads, loss and optimizer of the training step.
//   tdeed_dropout_fwd / _bwd   nn.Dropout(p) in front of the FC heads (model/modules.py:366-387), counter-based RNG
//   tdeed_linear_fwd / _bwd_data   small-N linear layers (K+1 class logits, 1 displacement output)
//   tdeed_ce_mse_loss          weighted cross entropy (hard int64 or soft mixup targets) + MSE on the displacement
//                              (model/model.py:208-211,308-319): loss value and d(loss)/d(logits, displ) in one launch
//   tdeed_adamw_step           torch.optim.AdamW (model/modules.py:37-39), fp32 state, optional bf16 shadow of the weights
//   tdeed_axpy                 y += alpha * x  (gradient accumulation, acc_grad_iter > 1)
#include <cmath>
#include "common.cuh"

namespace tdeed {

__device__ inline uint32_t hash_u64(unsigned long long z) {   // splitmix64 finaliser
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}

__global__ void dropout_fwd_kernel(const float* __restrict__ x, long long n, float p, unsigned long long seed,
                                   float* __restrict__ out, uint8_t* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = (float)hash_u64(seed * 0x100000001b3ull + (unsigned long long)i) * (1.f / 4294967296.f);
  const bool keep = u >= p;
  mask[i] = keep ? 1 : 0;
  out[i] = keep ? x[i] / (1.f - p) : 0.f;
}

// seed read from device memory (base[0] + salt): lets a CUDA graph replay draw a fresh mask every step
__global__ void dropout_fwd_devseed_kernel(const float* __restrict__ x, long long n, float p, const long long* __restrict__ seed_dev,
                                           unsigned long long salt, float* __restrict__ out, uint8_t* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long seed = (unsigned long long)seed_dev[0] * 2ull + salt;
  const float u = (float)hash_u64(seed * 0x100000001b3ull + (unsigned long long)i) * (1.f / 4294967296.f);
  const bool keep = u >= p;
  mask[i] = keep ? 1 : 0;
  out[i] = keep ? x[i] / (1.f - p) : 0.f;
}

__global__ void dropout_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ mask, long long n, float p,
                                   const float* __restrict__ add, float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = mask[i] ? dy[i] / (1.f - p) : 0.f;
  if (add) v += add[i];
  dx[i] = v;
}

// out[m, j] = sum_c x[m, c] * W[j, c] + b[j];  warp per row, N <= 64
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ W, const float* __restrict__ b, int N,
                  float* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  for (int j = 0; j < N; ++j) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(x[(size_t)m * C + c], W[(size_t)j * C + c], a);
    a = warp_sum(a);
    if (lane == 0) out[(size_t)m * ldo + j] = a + b[j];
  }
}

// dx[m, c] = sum_j dout[m, j] * W[j, c]  (+ add)
__global__ void linear_bwd_data_kernel(const float* __restrict__ dout, int ldd, long long total, int C, const float* __restrict__ W,
                                       int N, const float* __restrict__ add, float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long m = i / C;
  float a = add ? add[i] : 0.f;
  for (int j = 0; j < N; ++j) a = fmaf(dout[m * ldd + j], W[(size_t)j * C + c], a);
  dx[i] = a;
}

// single CTA: rows are few (B*T <= a few thousand).  out[0] = total loss, out[1] = CE part, out[2] = MSE part
__global__ void __launch_bounds__(256)
ce_mse_loss_kernel(const float* __restrict__ logits, int M, int K, int ld, const long long* __restrict__ hard,
                   const float* __restrict__ soft, const float* __restrict__ cw, const float* __restrict__ displ,
                   const float* __restrict__ labelD, float* __restrict__ out, float* __restrict__ dlogits,
                   float* __restrict__ ddispl) {
  __shared__ float s_red[32];
  float lsum = 0.f, wsum = 0.f, msum = 0.f;
  int bad = 0;                                   // a hard label outside [0, K): F.cross_entropy raises; here the loss becomes NaN
  for (int m = threadIdx.x; m < M; m += 256) {
    const float* z = logits + (size_t)m * ld;
    float mx = z[0];
    for (int j = 1; j < K; ++j) mx = fmaxf(mx, z[j]);
    float se = 0.f;
    for (int j = 0; j < K; ++j) se += expf(z[j] - mx);
    const float lse = mx + logf(se);
    if (hard) {
      const long long yl = hard[m];
      bad |= (yl < 0 || yl >= K) ? 1 : 0;
      const int y = (int)min(max(yl, 0ll), (long long)K - 1);       // never index out of bounds
      const float w = cw ? cw[y] : 1.f;
      lsum += w * (lse - z[y]);
      wsum += w;
    } else {
      float a = 0.f;
      for (int j = 0; j < K; ++j) a = fmaf((cw ? cw[j] : 1.f) * soft[(size_t)m * K + j], lse - z[j], a);
      lsum += a;
    }
    if (displ) {
      const float d = displ[m] - labelD[m];
      msum = fmaf(d, d, msum);
    }
  }
  lsum = block_sum(lsum, s_red);
  wsum = block_sum(wsum, s_red);
  msum = block_sum(msum, s_red);
  const float denom = hard ? wsum : (float)M;
  const int any_bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    const float ce = any_bad ? __int_as_float(0x7fc00000) : lsum / denom, mse = displ ? msum / (float)M : 0.f;
    out[0] = ce + mse;
    out[1] = ce;
    out[2] = mse;
  }
  for (int m = threadIdx.x; m < M; m += 256) {
    const float* z = logits + (size_t)m * ld;
    float mx = z[0];
    for (int j = 1; j < K; ++j) mx = fmaxf(mx, z[j]);
    float se = 0.f;
    for (int j = 0; j < K; ++j) se += expf(z[j] - mx);
    const float inv = 1.f / se;
    if (hard) {
      const int y = (int)min(max(hard[m], 0ll), (long long)K - 1);
      const float w = (cw ? cw[y] : 1.f) / denom;
      for (int j = 0; j < K; ++j) dlogits[(size_t)m * K + j] = w * (expf(z[j] - mx) * inv - (j == y ? 1.f : 0.f));
    } else {
      float tw = 0.f;
      for (int j = 0; j < K; ++j) tw = fmaf(cw ? cw[j] : 1.f, soft[(size_t)m * K + j], tw);
      for (int j = 0; j < K; ++j)
        dlogits[(size_t)m * K + j] = (expf(z[j] - mx) * inv * tw - (cw ? cw[j] : 1.f) * soft[(size_t)m * K + j]) / denom;
    }
    if (displ) ddispl[m] = 2.f * (displ[m] - labelD[m]) / (float)M;
  }
}

// Joint-dataset ("double head") loss of model/model.py:278-306: sample b uses head 1 (columns [0, n1)) when dataset[b] == 1 and
// head 2 (columns [n1, n1+n2)) when dataset[b] == 2; per sample F.cross_entropy(weight = cw[:k]) over its T rows (weighted mean
// for int64 targets — already shifted by n1 for dataset 2, as update_labels_2heads does — mean over rows for soft targets),
// summed over samples and divided by B; + MSE on the displacement over all rows.  Single CTA.
__global__ void __launch_bounds__(256)
ce_mse_loss_2h_kernel(const float* __restrict__ logits, int B, int T, int n1, int n2, int ld, const int* __restrict__ dataset,
                      const long long* __restrict__ hard, const float* __restrict__ soft, const float* __restrict__ cw,
                      const float* __restrict__ displ, const float* __restrict__ labelD, float* __restrict__ out,
                      float* __restrict__ dlogits, float* __restrict__ ddispl) {
  __shared__ float s_red[32];
  const int K = n1 + n2, M = B * T;
  float total = 0.f;
  int bad = 0;                                   // a hard label outside its head's range: the loss becomes NaN (torch raises)
  for (int b = 0; b < B; ++b) {
    const int ds = dataset[b];
    const int c0 = ds == 2 ? n1 : 0, k = ds == 2 ? n2 : n1;
    float lsum = 0.f, wsum = 0.f;
    if (ds == 1 || ds == 2) {
      for (int t = threadIdx.x; t < T; t += 256) {
        const int m = b * T + t;
        const float* z = logits + (size_t)m * ld + c0;
        float mx = z[0];
        for (int j = 1; j < k; ++j) mx = fmaxf(mx, z[j]);
        float se = 0.f;
        for (int j = 0; j < k; ++j) se += expf(z[j] - mx);
        const float lse = mx + logf(se);
        if (hard) {
          const long long yl = hard[m] - c0;
          bad |= (yl < 0 || yl >= k) ? 1 : 0;
          const int y = (int)min(max(yl, 0ll), (long long)k - 1);
          const float w = cw ? cw[y] : 1.f;
          lsum += w * (lse - z[y]);
          wsum += w;
        } else {
          float a = 0.f;
          for (int j = 0; j < k; ++j) a = fmaf((cw ? cw[j] : 1.f) * soft[(size_t)m * K + c0 + j], lse - z[j], a);
          lsum += a;
        }
      }
    }
    lsum = block_sum(lsum, s_red);
    wsum = block_sum(wsum, s_red);
    const float denom = (hard ? wsum : (float)T) * (float)B;
    if (ds == 1 || ds == 2) total += lsum / denom;
    for (int t = threadIdx.x; t < T; t += 256) {
      const int m = b * T + t;
      float* dz = dlogits + (size_t)m * K;
      for (int j = 0; j < K; ++j) dz[j] = 0.f;
      if (ds != 1 && ds != 2) continue;
      const float* z = logits + (size_t)m * ld + c0;
      float mx = z[0];
      for (int j = 1; j < k; ++j) mx = fmaxf(mx, z[j]);
      float se = 0.f;
      for (int j = 0; j < k; ++j) se += expf(z[j] - mx);
      const float inv = 1.f / se;
      if (hard) {
        const int y = (int)min(max(hard[m] - c0, 0ll), (long long)k - 1);
        const float w = (cw ? cw[y] : 1.f) / denom;
        for (int j = 0; j < k; ++j) dz[c0 + j] = w * (expf(z[j] - mx) * inv - (j == y ? 1.f : 0.f));
      } else {
        float tw = 0.f;
        for (int j = 0; j < k; ++j) tw = fmaf(cw ? cw[j] : 1.f, soft[(size_t)m * K + c0 + j], tw);
        for (int j = 0; j < k; ++j)
          dz[c0 + j] = (expf(z[j] - mx) * inv * tw - (cw ? cw[j] : 1.f) * soft[(size_t)m * K + c0 + j]) / denom;
      }
    }
  }
  float msum = 0.f;
  if (displ) {
    for (int m = threadIdx.x; m < M; m += 256) {
      const float d = displ[m] - labelD[m];
      msum = fmaf(d, d, msum);
      ddispl[m] = 2.f * d / (float)M;
    }
  }
  msum = block_sum(msum, s_red);
  const int any_bad = __syncthreads_or(bad);
  if (any_bad) total = __int_as_float(0x7fc00000);
  if (threadIdx.x == 0) {
    const float mse = displ ? msum / (float)M : 0.f;
    out[0] = total + mse;
    out[1] = total;
    out[2] = mse;
  }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                             float grad_scale, __nv_bfloat16* __restrict__ shadow) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = g[i] * grad_scale;
  float pv = p[i] * (1.f - lr * wd);
  const float mv = m[i] + (gr - m[i]) * (1.f - beta1);        // lerp_, as torch
  const float vv = v[i] * beta2 + (1.f - beta2) * gr * gr;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  pv -= (lr / bc1) * (mv / denom);
  p[i] = pv;
  m[i] = mv;
  v[i] = vv;
  if (shadow) shadow[i] = __float2bfloat16_rn(pv);
}

__global__ void axpy_kernel(const float* __restrict__ x, float alpha, long long n, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(alpha, x[i], y[i]);
}

}  // namespace tdeed

using namespace tdeed;

extern "C" int tdeed_dropout_fwd(const float* x, long long n, float p, unsigned long long seed, float* out, unsigned char* mask,
                                 void* stream) {
  TDEED_REQUIRE(x && out && mask && n > 0 && p >= 0.f && p < 1.f, TDEED_ERR_SHAPE, "tdeed_dropout_fwd: bad arguments");
  dropout_fwd_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, p, seed, out, mask);
  return check_launch("tdeed_dropout_fwd");
}

extern "C" int tdeed_dropout_fwd_devseed(const float* x, long long n, float p, const long long* seed_dev, unsigned long long salt,
                                         float* out, unsigned char* mask, void* stream) {
  TDEED_REQUIRE(x && out && mask && seed_dev && n > 0 && p >= 0.f && p < 1.f, TDEED_ERR_SHAPE, "tdeed_dropout_fwd_devseed: bad arguments");
  dropout_fwd_devseed_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, p, seed_dev, salt, out, mask);
  return check_launch("tdeed_dropout_fwd_devseed");
}

extern "C" int tdeed_dropout_bwd(const float* dy, const unsigned char* mask, long long n, float p, const float* add, float* dx,
                                 void* stream) {
  TDEED_REQUIRE(dy && mask && dx && n > 0 && p >= 0.f && p < 1.f, TDEED_ERR_SHAPE, "tdeed_dropout_bwd: bad arguments");
  dropout_bwd_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, mask, n, p, add, dx);
  return check_launch("tdeed_dropout_bwd");
}

extern "C" int tdeed_linear_fwd(const float* x, int M, int C, const float* W, const float* b, int N, float* out, int ldo,
                                void* stream) {
  TDEED_REQUIRE(x && W && b && out && M > 0 && C > 0 && N > 0 && N <= 64 && ldo >= N, TDEED_ERR_SHAPE, "tdeed_linear_fwd: bad arguments");
  linear_fwd_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(x, M, C, W, b, N, out, ldo);
  return check_launch("tdeed_linear_fwd");
}

extern "C" int tdeed_linear_bwd_data(const float* dout, int ldd, int M, int C, const float* W, int N, const float* add, float* dx,
                                     void* stream) {
  TDEED_REQUIRE(dout && W && dx && M > 0 && C > 0 && N > 0 && ldd >= N, TDEED_ERR_SHAPE, "tdeed_linear_bwd_data: bad arguments");
  const long long total = (long long)M * C;
  linear_bwd_data_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(dout, ldd, total, C, W, N, add, dx);
  return check_launch("tdeed_linear_bwd_data");
}

extern "C" int tdeed_ce_mse_loss(const float* logits, int M, int K, int ld_logits, const long long* target_hard,
                                 const float* target_soft, const float* class_weight, const float* displ, const float* labelD,
                                 float* loss_out, float* dlogits, float* ddispl, void* stream) {
  TDEED_REQUIRE(logits && loss_out && dlogits && M > 0 && K > 0 && K <= 1024 && ld_logits >= K, TDEED_ERR_SHAPE, "tdeed_ce_mse_loss: bad arguments");
  TDEED_REQUIRE((target_hard != nullptr) != (target_soft != nullptr), TDEED_ERR_SHAPE, "tdeed_ce_mse_loss: exactly one of hard / soft targets");
  TDEED_REQUIRE(!displ || (labelD && ddispl), TDEED_ERR_SHAPE, "tdeed_ce_mse_loss: displacement needs labelD and ddispl");
  ce_mse_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, M, K, ld_logits, target_hard, target_soft, class_weight, displ,
                                                          labelD, loss_out, dlogits, ddispl);
  return check_launch("tdeed_ce_mse_loss");
}

extern "C" int tdeed_ce_mse_loss_2heads(const float* logits, int B, int T, int n1, int n2, int ld_logits, const int* dataset,
                                        const long long* target_hard, const float* target_soft, const float* class_weight,
                                        const float* displ, const float* labelD, float* loss_out, float* dlogits, float* ddispl,
                                        void* stream) {
  TDEED_REQUIRE(logits && dataset && loss_out && dlogits && B > 0 && T > 0 && n1 > 0 && n2 > 0 && ld_logits >= n1 + n2, TDEED_ERR_SHAPE,
                "tdeed_ce_mse_loss_2heads: bad arguments");
  TDEED_REQUIRE((target_hard != nullptr) != (target_soft != nullptr), TDEED_ERR_SHAPE, "tdeed_ce_mse_loss_2heads: exactly one of hard / soft targets");
  TDEED_REQUIRE(!displ || (labelD && ddispl), TDEED_ERR_SHAPE, "tdeed_ce_mse_loss_2heads: displacement needs labelD and ddispl");
  ce_mse_loss_2h_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, B, T, n1, n2, ld_logits, dataset, target_hard, target_soft,
                                                             class_weight, displ, labelD, loss_out, dlogits, ddispl);
  return check_launch("tdeed_ce_mse_loss_2heads");
}

extern "C" int tdeed_adamw_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                                double eps, double weight_decay, int step, float grad_scale, void* shadow_bf16, void* stream) {
  TDEED_REQUIRE(p && g && m && v && n > 0 && step >= 1, TDEED_ERR_SHAPE, "tdeed_adamw_step: bad arguments");
  const float bc1 = (float)(1.0 - pow(beta1, (double)step));          // bias corrections in double on the host, like torch
  const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  adamw_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)lr, (float)beta1, (float)beta2,
                                                                                 (float)eps, (float)weight_decay, bc1, bc2_sqrt,
                                                                                 grad_scale, (__nv_bfloat16*)shadow_bf16);
  return check_launch("tdeed_adamw_step");
}

extern "C" int tdeed_axpy(const float* x, float alpha, long long n, float* y, void* stream) {
  TDEED_REQUIRE(x && y && n > 0, TDEED_ERR_SHAPE, "tdeed_axpy: bad arguments");
  axpy_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(x, alpha, n, y);
  return check_launch("tdeed_axpy");
}
