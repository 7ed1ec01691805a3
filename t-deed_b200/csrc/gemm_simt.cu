// (2a) CUDA-core GEMM with fused epilogue — the exact-fp32 backend of tdeed_gemm_fwd (parity mode:
// fp32 FMA accumulation, no tensor cores) and the gather (strided 1x1 conv) path.
//   out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] + residual[m, n] )
// 64x64 output tile per CTA, BK = 16, 256 threads, 4x4 register micro-tile per thread; operands are
// staged k-major in shared memory so the inner product reads are conflict-free float4s.
#include "common.cuh"

namespace tdeed {

struct SimtSegs {
  const void* a[TDEED_GEMM_MAX_SEGS];
  long long lda[TDEED_GEMM_MAX_SEGS];
  int col0[TDEED_GEMM_MAX_SEGS];
  int k[TDEED_GEMM_MAX_SEGS];
  int nseg;
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__device__ inline void ld4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ inline void ld4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}

__device__ inline float ld_any(const void* p, int dtype, size_t i) {
  return dtype == TDEED_F32 ? reinterpret_cast<const float*>(p)[i]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ inline void st_any(void* p, int dtype, size_t i, float v) {
  if (dtype == TDEED_F32) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(SimtSegs segs, long long M, int N, int K, int gstride, int gh, int gw,
                 const T* __restrict__ W, const float* __restrict__ bias,
                 const void* __restrict__ residual, long long ldr, int res_dtype, int act,
                 void* __restrict__ out, long long ldo, int out_dtype) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;

  // source row of this thread's A loads (constant over the k loop)
  long long arow = m0 + lrow;
  const bool arow_ok = arow < M;
  if (arow_ok && gstride > 1) {
    const int ow = (gw + gstride - 1) / gstride, oh = (gh + gstride - 1) / gstride;
    const long long f = arow / ((long long)oh * ow);
    const int rem = (int)(arow - f * (long long)oh * ow);
    const int oy = rem / ow, ox = rem - oy * ow;
    arow = (f * gh + (long long)oy * gstride) * gw + (long long)ox * gstride;
  }
  const int wrow = n0 + lrow;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    const int kg = k0 + lk;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
    if (kg < K) {
      if (arow_ok) {
        int s = 0, kk = kg;
        if (segs.nseg > 1 && kg >= segs.k[0]) { s = 1; kk = kg - segs.k[0]; }
        ld4(reinterpret_cast<const T*>(segs.a[s]) + arow * segs.lda[s] + segs.col0[s] + kk, av);
      }
      if (wrow < N) ld4(W + (size_t)wrow * K + kg, wv);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[lk + j][lrow] = av[j];
      Ws[lk + j][lrow] = wv[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (residual) v += ld_any(residual, res_dtype, (size_t)m * ldr + n);
      v = apply_act_rt(v, act);
      st_any(out, out_dtype, (size_t)m * ldo + n, v);
    }
  }
}

int gemm_simt_launch(int dtype, long long M, int N, int K, const SimtSegs& segs, int gstride, int gh, int gw,
                     const void* W, const float* bias, const void* residual, long long ldr, int res_dtype,
                     int act, void* out, long long ldo, int out_dtype, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div_ll(M, SG_BM), (unsigned)ceil_div(N, SG_BN));
  if (dtype == TDEED_F32)
    gemm_simt_kernel<float><<<grid, 256, 0, st>>>(segs, M, N, K, gstride, gh, gw, (const float*)W, bias, residual,
                                                  ldr, res_dtype, act, out, ldo, out_dtype);
  else
    gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(segs, M, N, K, gstride, gh, gw, (const __nv_bfloat16*)W,
                                                          bias, residual, ldr, res_dtype, act, out, ldo, out_dtype);
  return check_launch("tdeed_gemm_fwd(simt)");
}

}  // namespace tdeed
