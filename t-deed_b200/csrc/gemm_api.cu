// tdeed_gemm_fwd: backend selection between the exact CUDA-core kernel and the tcgen05 kernel.
#include "common.cuh"

namespace tdeed {
struct SimtSegs {
  const void* a[TDEED_GEMM_MAX_SEGS];
  long long lda[TDEED_GEMM_MAX_SEGS];
  int col0[TDEED_GEMM_MAX_SEGS];
  int k[TDEED_GEMM_MAX_SEGS];
  int nseg;
};
int gemm_simt_launch(int dtype, long long M, int N, int K, const SimtSegs& segs, int gstride, int gh, int gw,
                     const void* W, const float* bias, const void* residual, long long ldr, int res_dtype,
                     int act, void* out, long long ldo, int out_dtype, cudaStream_t st);
int gemm_tc_launch(long long M, int N, int K, int nseg, const tdeed_gemm_seg* segs, int gstride, int gh, int gw,
                   const void* W, const float* bias, const void* residual, long long ldr, int res_dtype, int act,
                   void* out, long long ldo, int out_dtype, cudaStream_t st, const float* a_scale = nullptr, int scale_rows = 0);
bool gemm_thin_applicable(int N, int K, int nseg, const tdeed_gemm_seg* segs, int gstride);
int gemm_thin_launch(long long M, int N, int K, int nseg, const tdeed_gemm_seg* segs, const void* W, const float* bias,
                     const void* residual, long long ldr, int res_dtype, int act, void* out, long long ldo, int out_dtype,
                     cudaStream_t st);
}  // namespace tdeed

extern "C" int tdeed_gemm_fwd(int dtype, long long M, int N, int nseg, const tdeed_gemm_seg* segs,
                              int gather_stride, int gather_h, int gather_w,
                              const void* W, const float* bias,
                              const void* residual, long long ldr, int res_dtype,
                              int act, void* out, long long ldo, int out_dtype, int backend, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(segs && W && out, TDEED_ERR_SHAPE, "tdeed_gemm_fwd: null pointer");
  TDEED_REQUIRE(nseg >= 1 && nseg <= TDEED_GEMM_MAX_SEGS, TDEED_ERR_SHAPE, "tdeed_gemm_fwd: nseg=%d", nseg);
  TDEED_REQUIRE(dtype == TDEED_F32 || dtype == TDEED_BF16, TDEED_ERR_UNSUPPORTED, "tdeed_gemm_fwd: dtype %d", dtype);
  TDEED_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && ldo % 8 == 0 && ldo >= N, TDEED_ERR_SHAPE,
                "tdeed_gemm_fwd: M=%lld N=%d ldo=%lld (N, ldo must be multiples of 8)", M, N, ldo);
  TDEED_REQUIRE(!residual || (ldr % 8 == 0 && ldr >= N), TDEED_ERR_SHAPE, "tdeed_gemm_fwd: ldr=%lld", ldr);
  int K = 0;
  SimtSegs ss{};
  ss.nseg = nseg;
  for (int s = 0; s < nseg; ++s) {
    TDEED_REQUIRE(segs[s].a && segs[s].k > 0 && segs[s].k % 4 == 0 && segs[s].col0 % 4 == 0 &&
                  segs[s].lda >= segs[s].col0 + segs[s].k, TDEED_ERR_SHAPE,
                  "tdeed_gemm_fwd: segment %d (col0=%d k=%d lda=%lld) must be 4-aligned and in range", s,
                  segs[s].col0, segs[s].k, segs[s].lda);
    ss.a[s] = segs[s].a; ss.lda[s] = segs[s].lda; ss.col0[s] = segs[s].col0; ss.k[s] = segs[s].k;
    K += segs[s].k;
  }
  TDEED_REQUIRE(gather_stride <= 1 || (gather_h > 0 && gather_w > 0 && nseg == 1), TDEED_ERR_SHAPE,
                "tdeed_gemm_fwd: gather needs one segment and a geometry");
  cudaStream_t st = (cudaStream_t)stream;
  bool aligned8 = K % 8 == 0;
  for (int s = 0; s < nseg; ++s) aligned8 = aligned8 && segs[s].col0 % 8 == 0 && (s == nseg - 1 || segs[s].k % 8 == 0);
  const bool thin_ok = dtype == TDEED_BF16 && gemm_thin_applicable(N, K, nseg, segs, gather_stride);
  if (backend == TDEED_GEMM_TCGEN05_THIN || (backend == TDEED_GEMM_AUTO && thin_ok)) {
    TDEED_REQUIRE(thin_ok, TDEED_ERR_UNSUPPORTED,
                  "tdeed_gemm_fwd: the thin-K backend needs bf16, K <= 64, N <= 256, 8-aligned segments and no gather");
    return gemm_thin_launch(M, N, K, nseg, segs, W, bias, residual, ldr, res_dtype, act, out, ldo, out_dtype, st);
  }
  bool use_tc = (backend == TDEED_GEMM_TCGEN05) || (backend == TDEED_GEMM_AUTO && dtype == TDEED_BF16 && aligned8);
  if (use_tc) {
    TDEED_REQUIRE(dtype == TDEED_BF16, TDEED_ERR_UNSUPPORTED, "tdeed_gemm_fwd: the tcgen05 backend needs bf16 operands");
    return gemm_tc_launch(M, N, K, nseg, segs, gather_stride, gather_h, gather_w, W, bias, residual, ldr, res_dtype, act,
                          out, ldo, out_dtype, st);
  }
  return gemm_simt_launch(dtype, M, N, K, ss, gather_stride, gather_h, gather_w, W, bias, residual, ldr, res_dtype,
                          act, out, ldo, out_dtype, st);
}

// (2c) conv3 of a RegNetY bottleneck with the squeeze-excite gate folded into the A operand (timm Bottleneck: x = conv3(se(x))):
//   out[m, n] = act( sum_k (A[m, k] * a_scale[m / scale_rows][k]) * W[n, k] + bias[n] + residual[m, n] )
// bf16 tcgen05 backend only (K > 64 layers: stages 3-4); A * gate is rounded to bf16 exactly like the stand-alone scale pass.
extern "C" int tdeed_gemm_scaled_fwd(long long M, int N, int K, const void* A, long long lda, const float* a_scale, int scale_rows,
                                     const void* W, const float* bias, const void* residual, long long ldr, int act, void* out,
                                     long long ldo, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(A && a_scale && W && out, TDEED_ERR_SHAPE, "tdeed_gemm_scaled_fwd: null pointer");
  TDEED_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && K > 0 && K % 8 == 0 && lda >= K && lda % 8 == 0 && ldo % 8 == 0 && ldo >= N &&
                scale_rows > 0 && M % scale_rows == 0, TDEED_ERR_SHAPE,
                "tdeed_gemm_scaled_fwd: M=%lld N=%d K=%d lda=%lld ldo=%lld scale_rows=%d", M, N, K, lda, ldo, scale_rows);
  TDEED_REQUIRE(!residual || (ldr % 8 == 0 && ldr >= N), TDEED_ERR_SHAPE, "tdeed_gemm_scaled_fwd: ldr=%lld", ldr);
  tdeed_gemm_seg seg{A, lda, 0, K};
  return gemm_tc_launch(M, N, K, 1, &seg, 1, 0, 0, W, bias, residual, ldr, TDEED_BF16, act, out, ldo, TDEED_BF16, (cudaStream_t)stream,
                        a_scale, scale_rows);
}
