// (10) classification + displacement heads, softmax and displacement scatter-max in one kernel.
// Reference: model/modules.py:366-387 (FCLayers/FC2Layers, dropout = identity in eval),
// model/model.py:141-146, model/modules.py:406-426 (process_prediction / process_double_head).
// Warp per (clip, frame) row: the row's features live in registers, each logit is a shuffle-reduced
// dot product; the scatter-max is order independent, so atomicMax on the (non-negative) float bits
// reproduces the reference's sequential Python loop exactly.
#include "common.cuh"

namespace tdeed {

constexpr int HD_THREADS = 256;
constexpr int HD_MAX_PER_LANE = 32;   // C <= 1024

__global__ void __launch_bounds__(HD_THREADS)
heads_kernel(const float* __restrict__ feat, int rows, int T, int C, const float* __restrict__ w_cls,
             const float* __restrict__ b_cls, int k_out, const float* __restrict__ w_displ,
             const float* __restrict__ b_displ, int k_softmax, float* __restrict__ logits,
             float* __restrict__ displ, float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (HD_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / T, t = row - b * T;
  float f[HD_MAX_PER_LANE];
  const int per = (C + 31) / 32;
#pragma unroll
  for (int i = 0; i < HD_MAX_PER_LANE; ++i) {
    const int c = lane + 32 * i;
    f[i] = (i < per && c < C) ? feat[(size_t)row * C + c] : 0.f;
  }
  float lg0 = -INFINITY, lg1 = -INFINITY;   // logit k lives in lane k%32, register k/32
  for (int k = 0; k < k_out; ++k) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HD_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      if (i < per && c < C) s = fmaf(f[i], w_cls[(size_t)k * C + c], s);
    }
    s = warp_sum(s) + b_cls[k];
    if (lane == (k & 31)) {
      if (k < 32) lg0 = s; else lg1 = s;
      logits[(size_t)row * k_out + k] = s;
    }
  }
  int dst_t = t;
  if (w_displ) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < HD_MAX_PER_LANE; ++i) {
      const int c = lane + 32 * i;
      if (i < per && c < C) s = fmaf(f[i], w_displ[c], s);
    }
    s = warp_sum(s) + b_displ[0];
    if (lane == 0) displ[row] = s;
    const float d = fminf(fmaxf(rintf(s), -1.0e6f), 1.0e6f);   // Tensor.round(): half to even
    dst_t = min(max(t - (int)d, 0), T - 1);
  }
  // softmax over the first k_softmax logits
  const float v0 = (lane < k_softmax) ? lg0 : -INFINITY;
  const float v1 = (lane + 32 < k_softmax) ? lg1 : -INFINITY;
  const float mx = warp_max(fmaxf(v0, v1));
  const float e0 = (lane < k_softmax) ? expf(v0 - mx) : 0.f;
  const float e1 = (lane + 32 < k_softmax) ? expf(v1 - mx) : 0.f;
  const float den = warp_sum(e0 + e1);
  float* prow = probs + ((size_t)b * T + dst_t) * k_softmax;
  if (w_displ) {
    if (lane < k_softmax) atomicMax(reinterpret_cast<int*>(prow + lane), __float_as_int(e0 / den));
    if (lane + 32 < k_softmax) atomicMax(reinterpret_cast<int*>(prow + lane + 32), __float_as_int(e1 / den));
  } else {
    if (lane < k_softmax) prow[lane] = e0 / den;
    if (lane + 32 < k_softmax) prow[lane + 32] = e1 / den;
  }
}

// softmax + scatter-max from precomputed logits / displacements (process_prediction as a standalone op)
__global__ void __launch_bounds__(HD_THREADS)
softmax_scatter_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ displ, int rows, int T,
                       int k_softmax, float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (HD_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / T, t = row - b * T;
  const float v0 = (lane < k_softmax) ? logits[(size_t)row * ld + lane] : -INFINITY;
  const float v1 = (lane + 32 < k_softmax) ? logits[(size_t)row * ld + lane + 32] : -INFINITY;
  const float mx = warp_max(fmaxf(v0, v1));
  const float e0 = (lane < k_softmax) ? expf(v0 - mx) : 0.f;
  const float e1 = (lane + 32 < k_softmax) ? expf(v1 - mx) : 0.f;
  const float den = warp_sum(e0 + e1);
  int dst_t = t;
  if (displ) {
    const float d = fminf(fmaxf(rintf(displ[row]), -1.0e6f), 1.0e6f);
    dst_t = min(max(t - (int)d, 0), T - 1);
  }
  float* prow = probs + ((size_t)b * T + dst_t) * k_softmax;
  if (displ) {
    if (lane < k_softmax) atomicMax(reinterpret_cast<int*>(prow + lane), __float_as_int(e0 / den));
    if (lane + 32 < k_softmax) atomicMax(reinterpret_cast<int*>(prow + lane + 32), __float_as_int(e1 / den));
  } else {
    if (lane < k_softmax) prow[lane] = e0 / den;
    if (lane + 32 < k_softmax) prow[lane + 32] = e1 / den;
  }
}

}  // namespace tdeed

extern "C" int tdeed_softmax_scatter_fwd(const float* logits, int ld_logits, const float* displ, int B, int T,
                                         int k_softmax, float* probs, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(logits && probs, TDEED_ERR_SHAPE, "tdeed_softmax_scatter_fwd: null pointer");
  TDEED_REQUIRE(B > 0 && T > 0 && k_softmax > 0 && k_softmax <= 64 && ld_logits >= k_softmax, TDEED_ERR_SHAPE,
                "tdeed_softmax_scatter_fwd: bad shape B=%d T=%d k=%d ld=%d", B, T, k_softmax, ld_logits);
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * T;
  if (displ) {
    cudaError_t e = cudaMemsetAsync(probs, 0, (size_t)rows * k_softmax * sizeof(float), st);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_softmax_scatter_fwd: memset: %s", cudaGetErrorString(e));
  }
  softmax_scatter_kernel<<<ceil_div(rows, HD_THREADS / 32), HD_THREADS, 0, st>>>(logits, ld_logits, displ, rows, T, k_softmax, probs);
  return check_launch("tdeed_softmax_scatter_fwd");
}

extern "C" int tdeed_heads_fwd(const float* feat, int B, int T, int C, const float* w_cls, const float* b_cls, int k_out,
                               const float* w_displ, const float* b_displ, int k_softmax,
                               float* logits, float* displ, float* probs, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(feat && w_cls && b_cls && logits && probs, TDEED_ERR_SHAPE, "tdeed_heads_fwd: null pointer");
  TDEED_REQUIRE(!w_displ || (b_displ && displ), TDEED_ERR_SHAPE, "tdeed_heads_fwd: displacement head needs bias and output");
  TDEED_REQUIRE(B > 0 && T > 0 && C > 0 && C <= 32 * HD_MAX_PER_LANE && k_out > 0 && k_out <= 64 && k_softmax > 0 &&
                k_softmax <= k_out, TDEED_ERR_SHAPE, "tdeed_heads_fwd: bad shape B=%d T=%d C=%d k_out=%d k_softmax=%d", B, T, C, k_out, k_softmax);
  cudaStream_t st = (cudaStream_t)stream;
  const int rows = B * T;
  if (w_displ) {
    cudaError_t e = cudaMemsetAsync(probs, 0, (size_t)rows * k_softmax * sizeof(float), st);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_heads_fwd: memset: %s", cudaGetErrorString(e));
  }
  heads_kernel<<<ceil_div(rows, HD_THREADS / 32), HD_THREADS, 0, st>>>(feat, rows, T, C, w_cls, b_cls, k_out, w_displ, b_displ,
                                                                       k_softmax, logits, displ, probs);
  return check_launch("tdeed_heads_fwd");
}
