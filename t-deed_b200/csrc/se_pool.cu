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

constexpr int SE_MAX_F = 4;    // frames per CTA (template parameter NF in {1, 2, 4})

// A CTA handles `fpc` consecutive frames so that the fc weights (2 * rd * c floats — more bytes than a 7x7 frame's
// activations) are fetched from L2 once per CTA instead of once per frame (r1c: the 7x7x368 SE launches were the
// slowest although their tensors are the smallest).
template <typename T, int NF>
__global__ void __launch_bounds__(SE_THREADS)
se_kernel(T* __restrict__ x, int n, int hw, int c, int rd, const float* __restrict__ w1,
          const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2) {
  extern __shared__ float smem[];
  const int c8n = c / 8;
  const int S = SE_THREADS / c8n > 0 ? SE_THREADS / c8n : 1;
  float* s_part = smem;                          // [S][c]
  float* s_mean = s_part + (size_t)S * c;        // [fpc][c]   (later reused as the scale)
  float* s_hid = s_mean + (size_t)NF * c;        // [NF][rd]
  const int f0 = blockIdx.x * NF;
  const int nf = min(NF, n - f0);
  const float inv = 1.f / (float)hw;

  for (int f = 0; f < nf; ++f) {
    frame_channel_sums(x + (size_t)(f0 + f) * hw * c, hw, c, s_part, s_mean + (size_t)f * c);
  }
  for (int i = threadIdx.x; i < nf * c; i += SE_THREADS) s_mean[i] *= inv;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rd; r += SE_THREADS / 32) {
    float acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
      const float wv = w1[(size_t)r * c + ch];
#pragma unroll
      for (int f = 0; f < NF; ++f)
        if (f < nf) acc[f] = fmaf(wv, s_mean[f * c + ch], acc[f]);
    }
    const float bb = b1[r];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      if (f < nf) {                                  // nf is CTA-uniform
        const float sres = warp_sum(acc[f]);
        if (lane == 0) s_hid[f * rd + r] = fmaxf(sres + bb, 0.f);
      }
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) {
    float acc[NF];
    const float bb = b2[ch];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = bb;
    for (int r = 0; r < rd; ++r) {
      const float wv = w2t[(size_t)r * c + ch];      // coalesced over ch
#pragma unroll
      for (int f = 0; f < NF; ++f)
        if (f < nf) acc[f] = fmaf(wv, s_hid[f * rd + r], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f)
      if (f < nf) s_mean[f * c + ch] = sigmoidf_(acc[f]);    // s_mean now holds the scale (each thread owns its channel)
  }
  __syncthreads();
  T* xf = x + (size_t)f0 * hw * c;
  const int per_frame = hw * c8n;
  for (int q = threadIdx.x; q < nf * per_frame; q += SE_THREADS) {
    const int f = q / per_frame, rem = q - f * per_frame;
    const int c8 = rem % c8n;
    float v[8];
    T* ptr = xf + (size_t)q * 8;                              // frames are contiguous: element offset = q * 8
    load8(ptr, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= s_mean[f * c + c8 * 8 + j];
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
  // frames per CTA: enough that a CTA's activation bytes outweigh the fc weights it has to fetch
  const size_t frame_bytes = (size_t)hw * c * (dtype == TDEED_BF16 ? 2 : 4);
  const size_t weight_bytes = (size_t)2 * rd * c * sizeof(float);
  int fpc = (int)((weight_bytes + frame_bytes - 1) / frame_bytes);
  // ... but never at the price of parallelism: the sums / scale passes are latency-bound streams, so keep at least
  // 8 CTAs per SM in flight (r1d: 8 frames per CTA at 3900 frames was 1.8x SLOWER than one frame per CTA)
  if (fpc > n / (8 * kNumSMs)) fpc = n / (8 * kNumSMs);
  fpc = fpc >= 4 ? 4 : (fpc >= 2 ? 2 : 1);
  const size_t smem = (part_floats(c) + (size_t)fpc * (c + rd)) * sizeof(float);
  TDEED_REQUIRE(smem <= 200 * 1024, TDEED_ERR_UNSUPPORTED, "tdeed_se_fwd: c=%d rd=%d need %zu B of shared memory", c, rd, smem);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(n, fpc);
  if (smem > 48 * 1024) fpc = 1;     // keep to the default shared-memory carve-out (c <= 2048 always fits with NF = 1)
  const size_t smem1 = (part_floats(c) + (size_t)fpc * (c + rd)) * sizeof(float);
#define TDEED_SE_LAUNCH(TT, NFV) se_kernel<TT, NFV><<<ceil_div(n, NFV), SE_THREADS, smem1, st>>>((TT*)x, n, hw, c, rd, w1, b1, w2, b2)
  if (dtype == TDEED_BF16) {
    if (fpc == 4) TDEED_SE_LAUNCH(__nv_bfloat16, 4); else if (fpc == 2) TDEED_SE_LAUNCH(__nv_bfloat16, 2); else TDEED_SE_LAUNCH(__nv_bfloat16, 1);
  } else if (dtype == TDEED_F32) {
    if (fpc == 4) TDEED_SE_LAUNCH(float, 4); else if (fpc == 2) TDEED_SE_LAUNCH(float, 2); else TDEED_SE_LAUNCH(float, 1);
  } else {
    set_error("tdeed_se_fwd: dtype %d", dtype);
    return TDEED_ERR_UNSUPPORTED;
  }
#undef TDEED_SE_LAUNCH
  (void)grid;
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
