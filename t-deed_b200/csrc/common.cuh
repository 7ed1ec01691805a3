// Shared helpers for the sm_100a kernels behind the C-ABI in include/tdeed_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../include/tdeed_b200.h"

namespace tdeed {

// thread-local last-error text, surfaced through tdeed_last_error()
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return TDEED_ERR_CUDA;
  }
  return TDEED_OK;
}

#define TDEED_REQUIRE(cond, code, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      ::tdeed::set_error(__VA_ARGS__);            \
      return (code);                              \
    }                                             \
  } while (0)

constexpr int kNumSMs = 148;

// Development knobs (TDEED_* environment variables: A/B switches, tracing, timing experiments that skip work) exist only
// in -DTDEED_DEV_KNOBS builds (`python t-deed_b200/build.py --dev`).  The release library never reads the environment.
#ifdef TDEED_DEV_KNOBS
inline const char* dev_env(const char* name) { return ::getenv(name); }
#else
inline const char* dev_env(const char*) { return nullptr; }
#endif

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- element access: kernels are templated on the activation type (float | __nv_bfloat16) ----
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kDtype = TDEED_F32;
  __device__ static float ld(const float* p) { return *p; }
  __device__ static void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kDtype = TDEED_BF16;
  __device__ static float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 8 consecutive elements <-> 8 floats (16 B for bf16, 32 B for f32); pointers must be 16 B aligned
__device__ inline void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ inline void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ inline void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ inline uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ inline void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 raw;
  raw.x = pack_bf16x2(v[0], v[1]);
  raw.y = pack_bf16x2(v[2], v[3]);
  raw.z = pack_bf16x2(v[4], v[5]);
  raw.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = raw;
}

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic block-wide sum (fixed tree); `scratch` holds >= 32 floats. All threads get the result.
__device__ inline float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = (lane < nwarp) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ inline float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ inline float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int ACT> __device__ inline float apply_act(float v) {
  if (ACT == TDEED_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == TDEED_ACT_GELU) return gelu_erf(v);
  return v;
}
__device__ inline float apply_act_rt(float v, int act) {
  if (act == TDEED_ACT_RELU) return fmaxf(v, 0.f);
  if (act == TDEED_ACT_GELU) return gelu_erf(v);
  return v;
}

}  // namespace tdeed
