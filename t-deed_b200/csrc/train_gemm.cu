// Weight-gradient GEMM and small reductions of the training step.
//   tdeed_gemm_tn   C[i, j] = sum_r A[r, i] * B[r, j]      (dW = dY^T X of every 1x1 conv / linear layer: the reduction
//                   runs over the M = frames*H*W rows, both operands are row-major with the reduction dim outermost)
//   tdeed_colsum    out[j] = sum_r X[r, j]                  (bias gradients)
//   tdeed_strided_add  dst[f, s*oy, s*ox, :] += src[f, oy, ox, :]   (data gradient of a stride-s 1x1 conv)
// CUDA-core fp32 accumulation; split over the rows with per-split partial tiles that are summed in a fixed order
// (deterministic, no atomics).
#include <cstdlib>
#include "common.cuh"

namespace tdeed {

constexpr int TN_TILE = 64, TN_KR = 16, TN_THREADS = 256;

template <typename TA, typename TB>
__global__ void __launch_bounds__(TN_THREADS)
gemm_tn_kernel(const TA* __restrict__ A, long long lda, const TB* __restrict__ B, long long ldb, long long R, int m, int n,
               int gather_stride, int gather_h, int gather_w, long long rows_per_split, float* __restrict__ out,
               long long ldo, long long split_stride) {
  __shared__ __align__(16) float sA[TN_KR][TN_TILE + 4];
  __shared__ __align__(16) float sB[TN_KR][TN_TILE + 4];
  const int i0 = blockIdx.y * TN_TILE, j0 = blockIdx.x * TN_TILE;
  const long long r_begin = (long long)blockIdx.z * rows_per_split;
  const long long r_end = min(R, r_begin + rows_per_split);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;      // thread computes rows i0+4*ty.., cols j0+4*tx..
  const int lrow = threadIdx.x / 16, lcol = (threadIdx.x % 16) * 4;
  const int goh = gather_stride > 1 ? (gather_h + gather_stride - 1) / gather_stride : 0;
  const int gow = gather_stride > 1 ? (gather_w + gather_stride - 1) / gather_stride : 0;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (long long r0 = r_begin; r0 < r_end; r0 += TN_KR) {
    const long long r = r0 + lrow;
    const bool rok = r < r_end;
    long long rb = r;
    if (gather_stride > 1 && rok) {
      const long long f = r / ((long long)goh * gow);
      const int rem = (int)(r - f * goh * gow);
      const int oy = rem / gow, ox = rem - oy * gow;
      rb = (f * gather_h + (long long)oy * gather_stride) * gather_w + (long long)ox * gather_stride;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ia = i0 + lcol + q, jb = j0 + lcol + q;
      sA[lrow][lcol + q] = (rok && ia < m) ? Elem<TA>::ld(A + r * lda + ia) : 0.f;
      sB[lrow][lcol + q] = (rok && jb < n) ? Elem<TB>::ld(B + rb * ldb + jb) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TN_KR; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[k][4 * ty]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[k][4 * tx]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* o = out + (size_t)blockIdx.z * split_stride;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + 4 * ty + a;
    if (i >= m) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + 4 * tx + b;
      if (j < n) o[(size_t)i * ldo + j] = acc[a][b];
    }
  }
}

__global__ void split_reduce_kernel(const float* __restrict__ part, int splits, long long split_stride, int m, int n,
                                    float* __restrict__ out, long long ldo, float alpha) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)m * n) return;
  const int i = (int)(idx / n), j = (int)(idx - (long long)i * n);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(size_t)z * split_stride + idx];
  out[(size_t)i * ldo + j] = alpha * s;
}

template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, long long M, int C, long long ld, long long rows_per_cta, float* __restrict__ part) {
  // thread per column (strided), serial over this CTA's rows
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += Elem<T>::ld(x + r * ld + c);
    part[(size_t)blockIdx.x * C + c] = s;
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ part, int nparts, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)part[(size_t)p * C + c];
  out[c] = (float)s;
}

template <typename T>
__global__ void __launch_bounds__(256)
strided_add_kernel(T* __restrict__ dst, const T* __restrict__ src, int h, int w, int c8n, int stride, int oh, int ow, long long total8) {
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  if (q >= total8) return;
  const int c8 = (int)(q % c8n);
  const long long p = q / c8n;
  const int ox = (int)(p % ow), oy = (int)((p / ow) % oh);
  const long long f = p / ((long long)ow * oh);
  T* d = dst + (((f * h + (long long)oy * stride) * w + (long long)ox * stride) * c8n + c8) * 8;
  float a[8], b[8];
  load8(d, a);
  load8(src + q * 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += b[j];
  store8(d, a);
}

// dst[f, oy, ox, :] = src[f, s*oy, s*ox, :]  (compact copy of the pixels a stride-s 1x1 conv reads: its weight gradient then
// runs on the dense tcgen05 dW GEMM)
template <typename T>
__global__ void __launch_bounds__(256)
strided_gather_kernel(const T* __restrict__ src, T* __restrict__ dst, int h, int w, int c8n, int stride, int oh, int ow, long long total8) {
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  if (q >= total8) return;
  const int c8 = (int)(q % c8n);
  const long long p = q / c8n;
  const int ox = (int)(p % ow), oy = (int)((p / ow) % oh);
  const long long f = p / ((long long)ow * oh);
  const T* s_ = src + (((f * h + (long long)oy * stride) * w + (long long)ox * stride) * c8n + c8) * 8;
  *reinterpret_cast<uint4*>(dst + q * 8) = *reinterpret_cast<const uint4*>(s_);
  if (sizeof(T) == 4) *reinterpret_cast<uint4*>(reinterpret_cast<char*>(dst + q * 8) + 16) = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(s_) + 16);
}

static int tn_splits(long long R, int m, int n) {
  const int tiles = ceil_div(m, TN_TILE) * ceil_div(n, TN_TILE);
  long long s = ceil_div_ll(2 * kNumSMs, tiles);                  // ~2 CTAs per SM
  const long long max_by_rows = ceil_div_ll(R, 8 * TN_KR);        // at least 128 rows per split
  if (s > max_by_rows) s = max_by_rows;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

template <typename TA, typename TB>
static int run_tn(const void* A, long long lda, const void* B, long long ldb, long long R, int m, int n, int gs, int gh, int gw,
                  float* out, long long ldo, float alpha, float* ws, cudaStream_t st) {
  const int splits = tn_splits(R, m, n);
  long long rps = ceil_div_ll(R, splits);
  rps = ceil_div_ll(rps, TN_KR) * TN_KR;
  dim3 grid(ceil_div(n, TN_TILE), ceil_div(m, TN_TILE), splits);
  const bool direct = splits == 1 && alpha == 1.f;
  gemm_tn_kernel<TA, TB><<<grid, TN_THREADS, 0, st>>>((const TA*)A, lda, (const TB*)B, ldb, R, m, n, gs, gh, gw, rps,
                                                      direct ? out : ws, direct ? ldo : n, (long long)m * n);
  int rc = check_launch("tdeed_gemm_tn");
  if (rc || direct) return rc;
  split_reduce_kernel<<<(unsigned)ceil_div_ll((long long)m * n, 256), 256, 0, st>>>(ws, splits, (long long)m * n, m, n, out, ldo, alpha);
  return check_launch("tdeed_gemm_tn(reduce)");
}

// tcgen05 backend (train_gemm_tc.cu): bf16 x bf16, no gather
bool gemm_tn_tc_applicable(int a_dtype, int b_dtype, const void* A, long long lda, const void* B, long long ldb, long long R,
                           int gather_stride);
long long gemm_tn_tc_workspace_floats(long long R, int m, int n);
int gemm_tn_tc_launch(const void* A, long long lda, const void* B, long long ldb, long long R, int m, int n, float alpha,
                      float* out, long long ldo, float* ws, cudaStream_t st);

static bool tn_force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = tdeed::dev_env("TDEED_GEMM_TN_SIMT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace tdeed

extern "C" long long tdeed_gemm_tn_workspace_floats(long long R, int m, int n) {
  const long long a = (long long)tdeed::tn_splits(R, m, n) * m * n;
  const long long b = tdeed::gemm_tn_tc_workspace_floats(R, m, n);
  return a > b ? a : b;
}

extern "C" int tdeed_gemm_tn(int a_dtype, const void* A, long long lda, int b_dtype, const void* B, long long ldb, long long R,
                             int m, int n, int gather_stride, int gather_h, int gather_w, float alpha, float* out,
                             long long ldo, float* workspace, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(A && B && out && workspace, TDEED_ERR_SHAPE, "tdeed_gemm_tn: null pointer");
  TDEED_REQUIRE(R > 0 && m > 0 && n > 0 && lda >= m && ldb >= n && ldo >= n && gather_stride >= 1, TDEED_ERR_SHAPE,
                "tdeed_gemm_tn: bad shape R=%lld m=%d n=%d", R, m, n);
  cudaStream_t st = (cudaStream_t)stream;
  if (!tn_force_simt() && gemm_tn_tc_applicable(a_dtype, b_dtype, A, lda, B, ldb, R, gather_stride))
    return gemm_tn_tc_launch(A, lda, B, ldb, R, m, n, alpha, out, ldo, workspace, st);
#define TN_CASE(DA, TA, DB, TB) \
  if (a_dtype == DA && b_dtype == DB) \
    return run_tn<TA, TB>(A, lda, B, ldb, R, m, n, gather_stride, gather_h, gather_w, out, ldo, alpha, workspace, st);
  TN_CASE(TDEED_F32, float, TDEED_F32, float)
  TN_CASE(TDEED_BF16, __nv_bfloat16, TDEED_BF16, __nv_bfloat16)
  TN_CASE(TDEED_F32, float, TDEED_BF16, __nv_bfloat16)
  TN_CASE(TDEED_BF16, __nv_bfloat16, TDEED_F32, float)
#undef TN_CASE
  set_error("tdeed_gemm_tn: dtypes %d/%d", a_dtype, b_dtype);
  return TDEED_ERR_UNSUPPORTED;
}

extern "C" long long tdeed_colsum_workspace_floats(long long M, int C) {
  return (long long)tdeed::kNumSMs * 2 * C;
}

extern "C" int tdeed_colsum(int dtype, const void* x, long long M, int C, long long ld, float* out, float* workspace, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && out && workspace && M > 0 && C > 0 && ld >= C, TDEED_ERR_SHAPE, "tdeed_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int nparts = (int)(M < 2 * kNumSMs ? M : 2 * kNumSMs);
  const long long rpc = ceil_div_ll(M, nparts);
  nparts = (int)ceil_div_ll(M, rpc);
  if (dtype == TDEED_BF16)
    colsum_partial_kernel<__nv_bfloat16><<<nparts, 256, 0, st>>>((const __nv_bfloat16*)x, M, C, ld, rpc, workspace);
  else if (dtype == TDEED_F32)
    colsum_partial_kernel<float><<<nparts, 256, 0, st>>>((const float*)x, M, C, ld, rpc, workspace);
  else { set_error("tdeed_colsum: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  int rc = check_launch("tdeed_colsum(partial)");
  if (rc) return rc;
  colsum_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(workspace, nparts, C, out);
  return check_launch("tdeed_colsum(final)");
}

extern "C" int tdeed_strided_add(int dtype, void* dst, const void* src, int n, int h, int w, int c, int stride, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(dst && src && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && stride >= 1, TDEED_ERR_SHAPE, "tdeed_strided_add: bad arguments");
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;
  const long long total8 = (long long)n * oh * ow * (c / 8);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div_ll(total8, 256);
  if (dtype == TDEED_BF16)
    strided_add_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, h, w, c / 8, stride, oh, ow, total8);
  else if (dtype == TDEED_F32)
    strided_add_kernel<float><<<grid, 256, 0, st>>>((float*)dst, (const float*)src, h, w, c / 8, stride, oh, ow, total8);
  else { set_error("tdeed_strided_add: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_strided_add");
}

extern "C" int tdeed_strided_gather(int dtype, const void* src, void* dst, int n, int h, int w, int c, int stride, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(dst && src && n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && stride >= 1, TDEED_ERR_SHAPE, "tdeed_strided_gather: bad arguments");
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;
  const long long total8 = (long long)n * oh * ow * (c / 8);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div_ll(total8, 256);
  if (dtype == TDEED_BF16)
    strided_gather_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, h, w, c / 8, stride, oh, ow, total8);
  else if (dtype == TDEED_F32)
    strided_gather_kernel<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, h, w, c / 8, stride, oh, ow, total8);
  else { set_error("tdeed_strided_gather: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_strided_gather");
}
