// (4) squeeze-excite in place and (6) global average pool + positional encoding.
// One CTA per frame; the frame's activations (<= a few hundred KB, just written by conv2) are read from
// L2.  All reductions use a fixed order -> bitwise deterministic.
#include "common.cuh"

namespace tdeed {

constexpr int SE_THREADS = 256;

// Per-channel sums over the hw pixels of one frame into s_sum[C] (deterministic).  s_part: [S][C] floats.
template <typename T>
__device__ inline void frame_channel_sums(const T* __restrict__ x, int hw, int c, float* s_part, float* s_sum) {
  const int c8n = c / 8;
  const int S = SE_THREADS / c8n > 0 ? SE_THREADS / c8n : 1;
  for (int q = threadIdx.x; q < c8n * S; q += SE_THREADS) {   // at most one pass when c8n <= 256
    const int c8 = q % c8n, seg = q / c8n;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = seg; p < hw; p += S) {
      float v[8];
      load8(x + (size_t)p * c + c8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[seg * c + c8 * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) {
    float s = 0.f;
    for (int seg = 0; seg < S; ++seg) s += s_part[seg * c + ch];
    s_sum[ch] = s;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(SE_THREADS)
se_kernel(T* __restrict__ x, int hw, int c, int rd, const float* __restrict__ w1, const float* __restrict__ b1,
          const float* __restrict__ w2t, const float* __restrict__ b2) {
  extern __shared__ float smem[];
  const int c8n = c / 8;
  const int S = SE_THREADS / c8n > 0 ? SE_THREADS / c8n : 1;
  float* s_part = smem;                 // [S][c]
  float* s_mean = s_part + (size_t)S * c;   // [c]
  float* s_hid = s_mean + c;            // [rd]
  float* s_scale = s_hid + rd;          // [c]
  T* xf = x + (size_t)blockIdx.x * hw * c;

  frame_channel_sums(xf, hw, c, s_part, s_mean);
  const float inv = 1.f / (float)hw;
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) s_mean[ch] *= inv;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rd; r += SE_THREADS / 32) {
    float s = 0.f;
    for (int ch = lane; ch < c; ch += 32) s = fmaf(w1[(size_t)r * c + ch], s_mean[ch], s);
    s = warp_sum(s);
    if (lane == 0) s_hid[r] = fmaxf(s + b1[r], 0.f);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) {
    float s = b2[ch];
    for (int r = 0; r < rd; ++r) s = fmaf(w2t[(size_t)r * c + ch], s_hid[r], s);   // coalesced over ch
    s_scale[ch] = sigmoidf_(s);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < hw * c8n; q += SE_THREADS) {
    const int c8 = q % c8n, p = q / c8n;
    float v[8];
    T* ptr = xf + (size_t)p * c + c8 * 8;
    load8(ptr, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= s_scale[c8 * 8 + j];
    store8(ptr, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(SE_THREADS)
pool_posenc_kernel(const T* __restrict__ x, int hw, int c, int clip_len, const float* __restrict__ temp_enc,
                   float* __restrict__ out) {
  extern __shared__ float smem[];
  const int c8n = c / 8;
  const int S = SE_THREADS / c8n > 0 ? SE_THREADS / c8n : 1;
  float* s_part = smem;
  float* s_sum = s_part + (size_t)S * c;
  const int f = blockIdx.x;
  frame_channel_sums(x + (size_t)f * hw * c, hw, c, s_part, s_sum);
  const float inv = 1.f / (float)hw;
  const float* te = temp_enc + (size_t)(f % clip_len) * c;
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) out[(size_t)f * c + ch] = s_sum[ch] * inv + te[ch];
}

static size_t part_floats(int c) {
  const int c8n = c / 8;
  const int S = SE_THREADS / c8n > 0 ? SE_THREADS / c8n : 1;
  return (size_t)S * c;
}

}  // namespace tdeed

extern "C" int tdeed_se_fwd(int dtype, void* x, int n, int hw, int c, int rd, const float* w1, const float* b1,
                            const float* w2, const float* b2, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && w1 && b1 && w2 && b2, TDEED_ERR_SHAPE, "tdeed_se_fwd: null pointer");
  TDEED_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 8 == 0 && c <= 2048 && rd > 0, TDEED_ERR_SHAPE,
                "tdeed_se_fwd: bad shape n=%d hw=%d c=%d rd=%d", n, hw, c, rd);
  const size_t smem = (part_floats(c) + 2 * (size_t)c + rd) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16)
    se_kernel<__nv_bfloat16><<<n, SE_THREADS, smem, st>>>((__nv_bfloat16*)x, hw, c, rd, w1, b1, w2, b2);
  else if (dtype == TDEED_F32)
    se_kernel<float><<<n, SE_THREADS, smem, st>>>((float*)x, hw, c, rd, w1, b1, w2, b2);
  else { set_error("tdeed_se_fwd: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_se_fwd");
}

extern "C" int tdeed_pool_posenc_fwd(int dtype, const void* x, int n, int hw, int c, int clip_len,
                                     const float* temp_enc, float* out, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && temp_enc && out, TDEED_ERR_SHAPE, "tdeed_pool_posenc_fwd: null pointer");
  TDEED_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 8 == 0 && c <= 2048 && clip_len > 0, TDEED_ERR_SHAPE,
                "tdeed_pool_posenc_fwd: bad shape n=%d hw=%d c=%d", n, hw, c);
  const size_t smem = (part_floats(c) + (size_t)c) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16)
    pool_posenc_kernel<__nv_bfloat16><<<n, SE_THREADS, smem, st>>>((const __nv_bfloat16*)x, hw, c, clip_len, temp_enc, out);
  else if (dtype == TDEED_F32)
    pool_posenc_kernel<float><<<n, SE_THREADS, smem, st>>>((const float*)x, hw, c, clip_len, temp_enc, out);
  else { set_error("tdeed_pool_posenc_fwd: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_pool_posenc_fwd");
}
