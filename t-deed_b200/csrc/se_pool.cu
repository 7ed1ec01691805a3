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

// Squeeze-excite as three kernels (r1e: the fused one-CTA-per-frame kernel fetched the 2*rd*c fc weights — 270 KB at
// 7x7x368, 7x the frame's own activations — once per frame and was the slowest op of every stage-4 block):
//   1. se_mean_kernel   per frame: channel means (fixed-order reduction)                       -> mean [n][c]
//   2. se_fc_kernel     per SE_FR frames: fc1 + ReLU + fc2 + sigmoid, weights fetched once per CTA -> scale [n][c]
//   3. se_scale_kernel  elementwise x *= scale[frame][channel], full-grid streaming pass
constexpr int SE_FR = 8;

template <typename T>
__global__ void __launch_bounds__(SE_THREADS)
se_mean_kernel(const T* __restrict__ x, int hw, int c, float* __restrict__ mean) {
  extern __shared__ float smem[];
  const int c8n = c / 8;
  const int S = SE_THREADS / c8n > 0 ? SE_THREADS / c8n : 1;
  float* s_part = smem;
  float* s_sum = s_part + (size_t)S * c;
  const int f = blockIdx.x;
  frame_channel_sums(x + (size_t)f * hw * c, hw, c, s_part, s_sum);
  const float inv = 1.f / (float)hw;
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) mean[(size_t)f * c + ch] = s_sum[ch] * inv;
}

// fc1 + ReLU + fc2 + sigmoid for F frames per CTA (F = 8, 16 or 32: about one wave of CTAs).  The fc weights (2*rd*c floats:
// 270 KB at stage 4 of RegNetY-200MF, 1.2 MB at 800MF) stay in L2 and are streamed once per CTA with coalesced loads; every
// weight meets all F frames of the CTA: F/4 conflict-free 16-byte shared-memory reads (s_mean is [F/4][c][4]) and F FMAs.
// The cross-lane sums of fc1 are a reduce-scatter (F - 1 shuffles per row instead of 5 F): lane l ends up with frame l % F.
// r1 processed 8 frames per CTA: one L2 round trip per 8 FMAs and 40 shuffles per row, 95 us per stage-4 launch.
template <int F>
__device__ __forceinline__ float reduce_scatter(float (&acc)[F], int lane) {
  // offsets >= F: plain butterfly (every lane keeps all F values); offsets < F: halve the set a lane is responsible for
#pragma unroll
  for (int off = 16; off >= F; off >>= 1)
#pragma unroll
    for (int j = 0; j < F; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
#pragma unroll
  for (int n = F / 2; n >= 1; n >>= 1) {
    const bool upper = (lane & n) != 0;
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float send = upper ? acc[j] : acc[j + n];
      const float keep = upper ? acc[j + n] : acc[j];
      acc[j] = keep + __shfl_xor_sync(0xffffffffu, send, n);
    }
  }
  return acc[0];                                 // sum over the warp for frame (lane % F)
}

constexpr int SE_FC_THREADS = 384;

template <int F>
__global__ void __launch_bounds__(SE_FC_THREADS, (F <= 16 ? 3 : 1))
se_fc_kernel(const float* __restrict__ mean, int n, int c, int rd, const float* __restrict__ w1,
             const float* __restrict__ b1, const float* __restrict__ w2t, const float* __restrict__ b2,
             float* __restrict__ scale) {
  extern __shared__ __align__(16) float smem[];
  constexpr int G = F / 4;
  float4* s_mean = reinterpret_cast<float4*>(smem);               // [G][c]  (4 frames per element)
  float* s_hid = smem + (size_t)F * c;                            // [rd][F]
  const int f0 = blockIdx.x * F;
  const int nf = min(F, n - f0);
  for (int i = threadIdx.x; i < F * c; i += SE_FC_THREADS) {
    const int f = i / c, ch = i - f * c;          // coalesced global read, scattered 4-byte store
    reinterpret_cast<float*>(s_mean + (size_t)(f >> 2) * c + ch)[f & 3] = (f < nf) ? mean[(size_t)(f0 + f) * c + ch] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rd; r += SE_FC_THREADS / 32) {
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
    const float* wr = w1 + (size_t)r * c;
#pragma unroll 4
    for (int ch = lane; ch < c; ch += 32) {        // 4 coalesced weight loads in flight
      const float wv = __ldg(wr + ch);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 mv = s_mean[(size_t)g * c + ch];
        acc[4 * g] = fmaf(wv, mv.x, acc[4 * g]);
        acc[4 * g + 1] = fmaf(wv, mv.y, acc[4 * g + 1]);
        acc[4 * g + 2] = fmaf(wv, mv.z, acc[4 * g + 2]);
        acc[4 * g + 3] = fmaf(wv, mv.w, acc[4 * g + 3]);
      }
    }
    const float sres = reduce_scatter<F>(acc, lane);
    if (lane < F) s_hid[r * F + lane] = fmaxf(sres + b1[r], 0.f);
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += SE_FC_THREADS) {
    float acc[F];
    const float bb = b2[ch];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = bb;
#pragma unroll 4
    for (int r = 0; r < rd; ++r) {
      const float wv = __ldg(w2t + (size_t)r * c + ch);      // coalesced over ch
      const float4* h = reinterpret_cast<const float4*>(s_hid + r * F);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 hv = h[g];                              // broadcast
        acc[4 * g] = fmaf(wv, hv.x, acc[4 * g]);
        acc[4 * g + 1] = fmaf(wv, hv.y, acc[4 * g + 1]);
        acc[4 * g + 2] = fmaf(wv, hv.z, acc[4 * g + 2]);
        acc[4 * g + 3] = fmaf(wv, hv.w, acc[4 * g + 3]);
      }
    }
#pragma unroll
    for (int f = 0; f < F; ++f)
      if (f < nf) scale[(size_t)(f0 + f) * c + ch] = sigmoidf_(acc[f]);
  }
}

template <int F>
static int launch_se_fc(const float* mean, int n, int c, int rd, const float* w1, const float* b1, const float* w2, const float* b2,
                        float* scale, cudaStream_t st) {
  const size_t smem_fc = (size_t)F * (c + rd) * sizeof(float);
  static bool set = false;
  if (smem_fc > 48 * 1024 && !set) {
    cudaError_t e = cudaFuncSetAttribute(se_fc_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_se_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    set = true;
  }
  se_fc_kernel<F><<<ceil_div(n, F), SE_FC_THREADS, smem_fc, st>>>(mean, n, c, rd, w1, b1, w2, b2, scale);
  return check_launch("tdeed_se_fwd(fc)");
}

// ---- fc1 + ReLU + fc2 + sigmoid on tensor cores (bf16 engine): 16 frames = the 16 rows of mma.sync m16n8k8 TF32 tiles ----
// The reference runs these two 1x1 convolutions under fp16 autocast (11-bit significands, fp32 accumulate); TF32 operands (11 bits,
// fp32 accumulate) keep that precision while the CUDA-core kernel above spends 2.7 instructions per FMA.  Means and hidden
// activations sit in shared memory (row stride = 4 mod 32 words: conflict-free fragment loads), the fp32 weights are read from
// L2 straight into B fragments (8 rows x 32 contiguous bytes per warp and k-step) and rounded with cvt.rna.tf32.
constexpr int SE_TC_THREADS = 384;      // 12 warps: rd = 92 is 12 n-tiles of fc1

__device__ __forceinline__ uint32_t se_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void se_mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__host__ __device__ inline int se_tc_stride(int k) { return k + ((4 - (k & 31)) & 31); }     // >= k, = 4 (mod 32)

__global__ void __launch_bounds__(SE_TC_THREADS, 3)
se_fc_tc_kernel(const float* __restrict__ mean, int n, int c, int rd, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2t, const float* __restrict__ b2, float* __restrict__ scale) {
  extern __shared__ __align__(16) uint32_t se_tc_smem[];
  const int rdp = (rd + 7) & ~7;
  const int sm_ld = se_tc_stride(c), sh_ld = se_tc_stride(rdp);
  uint32_t* s_mean = se_tc_smem;                       // [16][sm_ld]  tf32
  uint32_t* s_hid = s_mean + 16 * sm_ld;               // [16][sh_ld]  tf32
  const int f0 = blockIdx.x * 16;
  const int nf = min(16, n - f0);
  for (int i = threadIdx.x; i < 16 * c; i += SE_TC_THREADS) {
    const int f = i / c, ch = i - f * c;
    s_mean[f * sm_ld + ch] = se_tf32(f < nf ? mean[(size_t)(f0 + f) * c + ch] : 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  // fc1: hidden[16][rd] = relu(mean[16][c] . w1[rd][c]^T + b1)
  for (int n0 = warp * 8; n0 < rdp; n0 += (SE_TC_THREADS / 32) * 8) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool nok = n0 + g < rd;
    const float* wr = w1 + (size_t)(nok ? n0 + g : 0) * c + t;
    const uint32_t* ar = s_mean + g * sm_ld + t;
#pragma unroll 4
    for (int k0 = 0; k0 < c; k0 += 8) {
      const float w0 = nok ? __ldg(wr + k0) : 0.f, w4 = nok ? __ldg(wr + k0 + 4) : 0.f;
      se_mma_tf32(acc, ar[k0], ar[8 * sm_ld + k0], ar[k0 + 4], ar[8 * sm_ld + k0 + 4], se_tf32(w0), se_tf32(w4));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = n0 + 2 * t + j;
      const float bb = col < rd ? b1[col] : 0.f;
      s_hid[g * sh_ld + col] = col < rd ? se_tf32(fmaxf(acc[j] + bb, 0.f)) : 0u;
      s_hid[(g + 8) * sh_ld + col] = col < rd ? se_tf32(fmaxf(acc[2 + j] + bb, 0.f)) : 0u;
    }
  }
  __syncthreads();
  // fc2: scale[16][c] = sigmoid(hidden[16][rd] . w2t[rd][c] + b2)
  for (int n0 = warp * 8; n0 < c; n0 += (SE_TC_THREADS / 32) * 8) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wc = w2t + n0 + g;
    const uint32_t* ar = s_hid + g * sh_ld + t;
#pragma unroll 4
    for (int k0 = 0; k0 < rdp; k0 += 8) {
      const float w0 = k0 + t < rd ? __ldg(wc + (size_t)(k0 + t) * c) : 0.f;
      const float w4 = k0 + t + 4 < rd ? __ldg(wc + (size_t)(k0 + t + 4) * c) : 0.f;
      se_mma_tf32(acc, ar[k0], ar[8 * sh_ld + k0], ar[k0 + 4], ar[8 * sh_ld + k0 + 4], se_tf32(w0), se_tf32(w4));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = n0 + 2 * t + j;
      const float bb = b2[col];
      if (g < nf) scale[(size_t)(f0 + g) * c + col] = sigmoidf_(acc[j] + bb);
      if (g + 8 < nf) scale[(size_t)(f0 + g + 8) * c + col] = sigmoidf_(acc[2 + j] + bb);
    }
  }
}

static int launch_se_fc_tc(const float* mean, int n, int c, int rd, const float* w1, const float* b1, const float* w2, const float* b2,
                           float* scale, cudaStream_t st) {
  const int rdp = (rd + 7) & ~7;
  const size_t smem = (size_t)16 * (se_tc_stride(c) + se_tc_stride(rdp)) * sizeof(uint32_t);
  static size_t set = 48 * 1024;
  if (smem > set) {
    cudaError_t e = cudaFuncSetAttribute(se_fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_se_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    set = 200 * 1024;
  }
  se_fc_tc_kernel<<<ceil_div(n, 16), SE_TC_THREADS, smem, st>>>(mean, n, c, rd, w1, b1, w2, b2, scale);
  return check_launch("tdeed_se_fwd(fc_tc)");
}

// frames per CTA: about one wave of CTAs, bounded by shared memory (F * (c + rd) floats)
static int se_fc_dispatch(const float* mean, int n, int c, int rd, const float* w1, const float* b1, const float* w2, const float* b2,
                          float* scale, cudaStream_t st, bool tensor = false) {
  if (tensor && (size_t)16 * (tdeed::se_tc_stride(c) + tdeed::se_tc_stride((rd + 7) & ~7)) * 4 <= 190 * 1024)
    return launch_se_fc_tc(mean, n, c, rd, w1, b1, w2, b2, scale, st);
  // 16 frames per 512-thread CTA, two CTAs per SM: 32 warps per SM hide the L2 latency of the weight stream (one 8-warp CTA
  // of 32 frames per SM issued 28 % of the time: ncu r2); 8 frames when there are too few frames to fill the machine
  const int per = ceil_div(n, 2 * kNumSMs);
  int F = per > 8 ? 16 : 8;
  while (F > 8 && (size_t)F * (c + rd) * sizeof(float) > 190 * 1024) F >>= 1;
  if (F == 16) return launch_se_fc<16>(mean, n, c, rd, w1, b1, w2, b2, scale, st);
  return launch_se_fc<8>(mean, n, c, rd, w1, b1, w2, b2, scale, st);
}

template <typename T>
__global__ void __launch_bounds__(SE_THREADS)
se_scale_kernel(const T* x, T* out, long long total8, int per_frame8, int c8n, int c, const float* __restrict__ scale) {
  const long long q = (long long)blockIdx.x * SE_THREADS + threadIdx.x;
  if (q >= total8) return;
  const long long f = q / per_frame8;
  const int c8 = (int)(q % c8n);
  float v[8];
  load8(x + q * 8, v);
  const float4 s0 = *reinterpret_cast<const float4*>(scale + f * c + c8 * 8);
  const float4 s1 = *reinterpret_cast<const float4*>(scale + f * c + c8 * 8 + 4);
  v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
  v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
  store8(out + q * 8, v);
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

extern "C" long long tdeed_se_workspace_floats(int n, int c) { return 2LL * n * c; }

static int se_forward(int dtype, const void* x, void* out, int n, int hw, int c, int rd, const float* w1, const float* b1,
                      const float* w2, const float* b2, float* workspace, void* stream, bool apply = true, bool train = false) {
  using namespace tdeed;
  TDEED_REQUIRE(x && w1 && b1 && w2 && b2 && workspace, TDEED_ERR_SHAPE, "tdeed_se_fwd: null pointer");
  TDEED_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 8 == 0 && c <= 2048 && rd > 0 && rd <= 1024, TDEED_ERR_SHAPE,
                "tdeed_se_fwd: bad shape n=%d hw=%d c=%d rd=%d", n, hw, c, rd);
  TDEED_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, TDEED_ERR_SHAPE, "tdeed_se_fwd: workspace must be 16-byte aligned");
  float* mean = workspace;
  float* scale = workspace + (size_t)n * c;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_mean = (part_floats(c) + (size_t)c) * sizeof(float);
  const long long total8 = (long long)n * hw * (c / 8);
  const unsigned scale_grid = (unsigned)ceil_div_ll(total8, SE_THREADS);
  if (dtype == TDEED_BF16) {
    se_mean_kernel<__nv_bfloat16><<<n, SE_THREADS, smem_mean, st>>>((const __nv_bfloat16*)x, hw, c, mean);
  } else if (dtype == TDEED_F32) {
    se_mean_kernel<float><<<n, SE_THREADS, smem_mean, st>>>((const float*)x, hw, c, mean);
  } else {
    set_error("tdeed_se_fwd: dtype %d", dtype);
    return TDEED_ERR_UNSUPPORTED;
  }
  int rc = check_launch("tdeed_se_fwd(mean)");
  if (rc) return rc;
  rc = se_fc_dispatch(mean, n, c, rd, w1, b1, w2, b2, scale, st, dtype == TDEED_BF16 && !train);
  if (rc || !apply) return rc;
  if (dtype == TDEED_BF16)
    se_scale_kernel<__nv_bfloat16><<<scale_grid, SE_THREADS, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, total8, hw * (c / 8), c / 8, c, scale);
  else
    se_scale_kernel<float><<<scale_grid, SE_THREADS, 0, st>>>((const float*)x, (float*)out, total8, hw * (c / 8), c / 8, c, scale);
  return check_launch("tdeed_se_fwd(scale)");
}

extern "C" int tdeed_se_fwd(int dtype, void* x, int n, int hw, int c, int rd, const float* w1, const float* b1,
                            const float* w2, const float* b2, float* workspace, void* stream) {
  return se_forward(dtype, x, x, n, hw, c, rd, w1, b1, w2, b2, workspace, stream);
}

// gate only: mean + fc; the per-(frame, channel) gate lands in workspace[n*c, 2*n*c) and is applied by the consumer
// (tdeed_gemm_scaled_fwd folds it into conv3's A operand: no read-modify-write pass over the activation)
extern "C" int tdeed_se_gate_fwd(int dtype, const void* x, int n, int hw, int c, int rd, const float* w1, const float* b1,
                                 const float* w2, const float* b2, float* workspace, void* stream) {
  return se_forward(dtype, x, nullptr, n, hw, c, rd, w1, b1, w2, b2, workspace, stream, false);
}

// training: out-of-place (the unscaled activation is needed by the backward pass); workspace keeps the per-frame means
// [n][c] and scales [n][c] for tdeed_se_bwd
extern "C" int tdeed_se_train_fwd(int dtype, const void* x, void* out, int n, int hw, int c, int rd, const float* w1,
                                  const float* b1, const float* w2t, const float* b2, float* workspace, void* stream) {
  TDEED_REQUIRE(out, TDEED_ERR_SHAPE, "tdeed_se_train_fwd: null pointer");
  return se_forward(dtype, x, out, n, hw, c, rd, w1, b1, w2t, b2, workspace, stream, true, true);   // exact fp32 fc: the backward differentiates it
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

// ---- squeeze-excite backward -------------------------------------------------------------------------------------
//   u = x * s[f, c]:   ds[f,c] = sum_hw du * x;   dv = ds * s (1 - s);   h = relu(W1 m + b1);
//   dh = (W2^T dv) * (h > 0);   dm = W1^T dh;   dx = du * s + dm / hw
// Weight gradients are left to tdeed_gemm_tn / tdeed_colsum on the per-frame vectors written to `vec`:
//   vec = dv [n][c] | dh [n][rd] | h [n][rd] | dm [n][c]
namespace tdeed {

template <typename T>
__global__ void __launch_bounds__(SE_THREADS)
se_bwd_ds_kernel(const T* __restrict__ x, const T* __restrict__ du, int hw, int c, float* __restrict__ ds) {
  extern __shared__ float smem[];
  const int c8n = c / 8;
  const int S = SE_THREADS / c8n > 0 ? SE_THREADS / c8n : 1;
  float* s_part = smem;
  const size_t base = (size_t)blockIdx.x * hw * c;
  for (int q = threadIdx.x; q < c8n * S; q += SE_THREADS) {
    const int c8 = q % c8n, seg = q / c8n;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = seg; p < hw; p += S) {
      float a[8], b[8];
      load8(x + base + (size_t)p * c + c8 * 8, a);
      load8(du + base + (size_t)p * c + c8 * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(a[j], b[j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[seg * c + c8 * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) {
    float s = 0.f;
    for (int seg = 0; seg < S; ++seg) s += s_part[seg * c + ch];
    ds[(size_t)blockIdx.x * c + ch] = s;
  }
}

// one CTA per frame (the fc matrices are small next to the activations in training batches)
__global__ void __launch_bounds__(SE_THREADS)
se_bwd_fc_kernel(const float* __restrict__ mean, const float* __restrict__ scale, const float* __restrict__ ds, int c, int rd,
                 const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2t,
                 float* __restrict__ dv_out, float* __restrict__ dh_out, float* __restrict__ h_out, float* __restrict__ dm_out) {
  extern __shared__ float smem[];
  float* s_m = smem;            // [c]
  float* s_dv = s_m + c;        // [c]
  float* s_h = s_dv + c;        // [rd]
  float* s_dh = s_h + rd;       // [rd]
  const int f = blockIdx.x;
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) {
    const float s = scale[(size_t)f * c + ch];
    const float dv = ds[(size_t)f * c + ch] * s * (1.f - s);
    s_m[ch] = mean[(size_t)f * c + ch];
    s_dv[ch] = dv;
    dv_out[(size_t)f * c + ch] = dv;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rd; r += SE_THREADS / 32) {
    float a = 0.f, g = 0.f;
    for (int ch = lane; ch < c; ch += 32) {
      a = fmaf(w1[(size_t)r * c + ch], s_m[ch], a);
      g = fmaf(w2t[(size_t)r * c + ch], s_dv[ch], g);
    }
    a = warp_sum(a);
    g = warp_sum(g);
    if (lane == 0) {
      const float hv = fmaxf(a + b1[r], 0.f);
      const float dh = hv > 0.f ? g : 0.f;
      s_h[r] = hv;
      s_dh[r] = dh;
      h_out[(size_t)f * rd + r] = hv;
      dh_out[(size_t)f * rd + r] = dh;
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += SE_THREADS) {
    float a = 0.f;
    for (int r = 0; r < rd; ++r) a = fmaf(w1[(size_t)r * c + ch], s_dh[r], a);
    dm_out[(size_t)f * c + ch] = a;
  }
}

template <typename T>
__global__ void __launch_bounds__(SE_THREADS)
se_bwd_dx_kernel(const T* __restrict__ du, long long total8, int per_frame8, int c8n, int c, float inv_hw,
                 const float* __restrict__ scale, const float* __restrict__ dm, T* __restrict__ dx) {
  const long long q = (long long)blockIdx.x * SE_THREADS + threadIdx.x;
  if (q >= total8) return;
  const long long f = q / per_frame8;
  const int c0 = (int)(q % c8n) * 8;
  float v[8];
  load8(du + q * 8, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], scale[f * c + c0 + j], dm[f * c + c0 + j] * inv_hw);
  store8(dx + q * 8, v);
}

// dz[f, p, :] = dfeat[f, :] / hw (backward of the global average pool), and d temp_enc[t, :] = sum_b dfeat[b*T + t, :]
template <typename T>
__global__ void __launch_bounds__(SE_THREADS)
pool_bwd_kernel(const float* __restrict__ dfeat, long long total8, int per_frame8, int c8n, int c, float inv_hw, T* __restrict__ dz) {
  const long long q = (long long)blockIdx.x * SE_THREADS + threadIdx.x;
  if (q >= total8) return;
  const long long f = q / per_frame8;
  const int c0 = (int)(q % c8n) * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = dfeat[f * c + c0 + j] * inv_hw;
  store8(dz + q * 8, v);
}

__global__ void temp_enc_bwd_kernel(const float* __restrict__ dfeat, int clips, int clip_len, int c, float* __restrict__ dte) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= clip_len * c) return;
  float s = 0.f;
  for (int b = 0; b < clips; ++b) s += dfeat[(size_t)b * clip_len * c + i];
  dte[i] = s;
}

}  // namespace tdeed

extern "C" long long tdeed_se_bwd_vec_floats(int n, int c, int rd) { return (long long)n * (3LL * c + 2LL * rd); }

extern "C" int tdeed_se_bwd(int dtype, const void* x, const void* du, int n, int hw, int c, int rd, const float* w1,
                            const float* b1, const float* w2t, const float* fwd_workspace, void* dx, float* vec, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && du && w1 && b1 && w2t && fwd_workspace && dx && vec, TDEED_ERR_SHAPE, "tdeed_se_bwd: null pointer");
  TDEED_REQUIRE(n > 0 && hw > 0 && c > 0 && c % 8 == 0 && c <= 2048 && rd > 0 && rd <= 1024, TDEED_ERR_SHAPE,
                "tdeed_se_bwd: bad shape n=%d hw=%d c=%d rd=%d", n, hw, c, rd);
  const float* mean = fwd_workspace;
  const float* scale = fwd_workspace + (size_t)n * c;
  float* dv = vec;
  float* dh = dv + (size_t)n * c;
  float* hh = dh + (size_t)n * rd;
  float* dm = hh + (size_t)n * rd;
  float* ds = dm + (size_t)n * c;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_ds = part_floats(c) * sizeof(float);
  const size_t smem_fc = (size_t)(2 * c + 2 * rd) * sizeof(float);
  const long long total8 = (long long)n * hw * (c / 8);
  const unsigned grid = (unsigned)ceil_div_ll(total8, SE_THREADS);
  if (dtype == TDEED_BF16)
    se_bwd_ds_kernel<__nv_bfloat16><<<n, SE_THREADS, smem_ds, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)du, hw, c, ds);
  else if (dtype == TDEED_F32)
    se_bwd_ds_kernel<float><<<n, SE_THREADS, smem_ds, st>>>((const float*)x, (const float*)du, hw, c, ds);
  else { set_error("tdeed_se_bwd: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  int rc = check_launch("tdeed_se_bwd(ds)");
  if (rc) return rc;
  se_bwd_fc_kernel<<<n, SE_THREADS, smem_fc, st>>>(mean, scale, ds, c, rd, w1, b1, w2t, dv, dh, hh, dm);
  rc = check_launch("tdeed_se_bwd(fc)");
  if (rc) return rc;
  if (dtype == TDEED_BF16)
    se_bwd_dx_kernel<__nv_bfloat16><<<grid, SE_THREADS, 0, st>>>((const __nv_bfloat16*)du, total8, hw * (c / 8), c / 8, c, 1.f / hw, scale, dm, (__nv_bfloat16*)dx);
  else
    se_bwd_dx_kernel<float><<<grid, SE_THREADS, 0, st>>>((const float*)du, total8, hw * (c / 8), c / 8, c, 1.f / hw, scale, dm, (float*)dx);
  return check_launch("tdeed_se_bwd(dx)");
}

extern "C" int tdeed_pool_posenc_bwd(int dtype, const float* dfeat, int clips, int clip_len, int hw, int c, void* dz,
                                     float* d_temp_enc, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(dfeat && dz && d_temp_enc, TDEED_ERR_SHAPE, "tdeed_pool_posenc_bwd: null pointer");
  TDEED_REQUIRE(clips > 0 && clip_len > 0 && hw > 0 && c > 0 && c % 8 == 0, TDEED_ERR_SHAPE, "tdeed_pool_posenc_bwd: bad shape");
  const long long n = (long long)clips * clip_len;
  const long long total8 = n * hw * (c / 8);
  const unsigned grid = (unsigned)ceil_div_ll(total8, SE_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16)
    pool_bwd_kernel<__nv_bfloat16><<<grid, SE_THREADS, 0, st>>>(dfeat, total8, hw * (c / 8), c / 8, c, 1.f / hw, (__nv_bfloat16*)dz);
  else if (dtype == TDEED_F32)
    pool_bwd_kernel<float><<<grid, SE_THREADS, 0, st>>>(dfeat, total8, hw * (c / 8), c / 8, c, 1.f / hw, (float*)dz);
  else { set_error("tdeed_pool_posenc_bwd: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  int rc = check_launch("tdeed_pool_posenc_bwd");
  if (rc) return rc;
  temp_enc_bwd_kernel<<<ceil_div(clip_len * c, 256), 256, 0, st>>>(dfeat, clips, clip_len, c, d_temp_enc);
  return check_launch("tdeed_pool_posenc_bwd(temp_enc)");
}
