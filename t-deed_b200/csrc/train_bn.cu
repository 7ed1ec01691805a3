// Training-mode BatchNorm around the convolutions (timm ConvNormAct / BatchNormAct2d in train(); the reference trains the
// whole backbone, model/model.py:202).  The conv kernels write the RAW convolution output y [M, C] (M = frames*H*W); then
//   tdeed_bn_stats     per-channel batch mean / biased variance (+ running-statistics update, momentum 0.1, unbiased var)
//   tdeed_bn_act_fwd   z = act(y*scale + shift (+ residual))
//   tdeed_bn_act_bwd   g = dz * (z > 0);  dgamma = sum g*xhat;  dbeta = sum g;
//                      dy = scale * (g - dbeta/M - xhat*dgamma/M);  optionally dres = g (shortcut branch)
// All per-channel reductions are two-stage with a fixed order (per-thread serial -> shared-memory tree in a fixed order ->
// per-CTA partials summed in double by one thread per channel): bitwise deterministic, no atomics.
#include "train_reduce.cuh"
#include <cstdlib>

namespace tdeed {

// ---- statistics: sums of (x - pivot) and (x - pivot)^2, pivot = first row (kills the E[x^2]-E[x]^2 cancellation) ----
template <typename T>
struct StatsOp {
  const T* x;
  long long ld;
  float piv[8];
  __device__ void begin(int ch0, int nch) { load_n(x + ch0, nch, piv); }
  static constexpr int kBatch = 8;
  struct Regs { float v[8]; };
  __device__ void load(long long r, int ch0, Regs& g) const { load8(x + r * ld + ch0, g.v); }
  __device__ void acc(const Regs& g, float (&a0)[8], float (&a1)[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = g.v[j] - piv[j];
      a0[j] += d;
      a1[j] = fmaf(d, d, a1[j]);
    }
  }
  __device__ void row(long long r, int ch0, int nch, float (&a0)[8], float (&a1)[8]) {
    Regs g;
    load_n(x + r * ld + ch0, nch, g.v);
    acc(g, a0, a1);
  }
};

template <typename T>
__global__ void bn_stats_final_kernel(const T* __restrict__ x, const float* __restrict__ part, int nparts, long long M, int C,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                      float* __restrict__ running_mean, float* __restrict__ running_var,
                                      float* __restrict__ stats) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp per channel
  if (ch >= C) return;
  const int cpad = ((C + 7) / 8) * 8;
  double s, q;
  warp_partial_sums(part, nparts, cpad, ch, s, q);
  if ((threadIdx.x & 31) != 0) return;
  const double piv = (double)Elem<T>::ld(x + ch);
  const double dm = s / (double)M;
  const double mean = piv + dm;
  double var = q / (double)M - dm * dm;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
  stats[ch] = (float)mean;
  stats[C + ch] = invstd;
  stats[2 * C + ch] = g * invstd;
  stats[3 * C + ch] = b - (float)mean * g * invstd;
  if (running_mean) {
    const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unb;
  }
}

// ---- forward apply.  A thread owns one channel octet (its scale / shift stay in registers) and walks the rows ----
template <typename T>
__global__ void __launch_bounds__(BN_THREADS)
bn_act_fwd_kernel(const T* __restrict__ y, long long M, int C, const float* __restrict__ scale,
                  const float* __restrict__ shift, const T* __restrict__ residual, int relu, T* __restrict__ out) {
  const BnLayout l = bn_layout(C);
  const int c8 = threadIdx.x % l.c8n, lr = threadIdx.x / l.c8n;
  if (lr >= l.rpi) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[c8 * 8 + j];
    sh[j] = shift[c8 * 8 + j];
  }
  for (long long r = (long long)blockIdx.x * l.rpi + lr; r < M; r += (long long)gridDim.x * l.rpi) {
    const long long at = r * C + c8 * 8;
    float v[8], rr[8];
    load8(y + at, v);
    if (residual) load8(residual + at, rr);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = fmaf(v[j], sc[j], sh[j]);
      if (residual) t += rr[j];
      v[j] = relu ? fmaxf(t, 0.f) : t;
    }
    store8(out + at, v);
  }
}

// ---- backward ----
// MODE: 0 no activation, 1 ReLU mask recomputed from y as (y*scale + shift > 0) (the z == y sentinel: saves reading z),
//       2 ReLU mask from the post-activation tensor z
template <typename T, int MODE, int KB = 4>
struct BwdOp {
  const T* dz;
  const T* z;
  const T* y;
  const float* mean;
  const float* invstd;
  int C;
  float mu[8], is[8], sc[8], sh[8];
  __device__ void begin(int ch0, int nch) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = j < nch ? mean[ch0 + j] : 0.f;
      is[j] = j < nch ? invstd[ch0 + j] : 0.f;
      sc[j] = j < nch ? mean[2 * C + ch0 + j] : 0.f;
      sh[j] = j < nch ? mean[3 * C + ch0 + j] : 0.f;
    }
  }
  static constexpr int kBatch = KB;
  struct Regs { float g[8], yv[8], zv[MODE == 2 ? 8 : 1]; };
  __device__ void load(long long r, int ch0, Regs& q) const {
    load8(dz + r * C + ch0, q.g);
    load8(y + r * C + ch0, q.yv);
    if constexpr (MODE == 2) load8(z + r * C + ch0, q.zv);
  }
  __device__ void acc(const Regs& q, float (&a0)[8], float (&a1)[8]) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bool keep = true;
      if constexpr (MODE == 1) keep = fmaf(q.yv[j], sc[j], sh[j]) > 0.f;
      if constexpr (MODE == 2) keep = q.zv[j] > 0.f;
      const float gg = keep ? q.g[j] : 0.f;
      a0[j] += gg;
      a1[j] = fmaf(gg, (q.yv[j] - mu[j]) * is[j], a1[j]);
    }
  }
  __device__ void row(long long r, int ch0, int nch, float (&a0)[8], float (&a1)[8]) {
    Regs q;
    load(r, ch0, q);
    acc(q, a0, a1);
  }
};

__global__ void bn_bwd_final_kernel(const float* __restrict__ part, int nparts, long long M, int C,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp per channel
  if (ch >= C) return;
  const int cpad = ((C + 7) / 8) * 8;
  double s, q;
  warp_partial_sums(part, nparts, cpad, ch, s, q);
  if ((threadIdx.x & 31) != 0) return;
  if (dbeta) dbeta[ch] = (float)s;
  if (dgamma) dgamma[ch] = (float)q;
  coef[ch] = (float)(s / (double)M);
  coef[C + ch] = (float)(q / (double)M);
}

template <typename T, int MODE, int U>
__global__ void __launch_bounds__(BN_THREADS)
bn_act_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ z, const T* __restrict__ y, long long M, int C,
                        const float* __restrict__ stats, const float* __restrict__ coef, T* __restrict__ dy,
                        T* __restrict__ dres) {
  const BnLayout l = bn_layout(C);
  const int c8 = threadIdx.x % l.c8n, lr = threadIdx.x / l.c8n;
  if (lr >= l.rpi) return;
  float mu[8], is[8], sc[8], sh[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = c8 * 8 + j;
    mu[j] = stats[ch];
    is[j] = stats[C + ch];
    sc[j] = stats[2 * C + ch];
    sh[j] = stats[3 * C + ch];
    k1[j] = coef[ch];
    k2[j] = coef[C + ch];
  }
  // U rows per trip, the loads of all of them issued before the first store
  const long long stride = (long long)gridDim.x * l.rpi;
  auto one = [&](long long at, float (&g)[8], const float (&yv)[8], const float* zv) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bool keep = true;
      if constexpr (MODE == 1) keep = fmaf(yv[j], sc[j], sh[j]) > 0.f;
      if constexpr (MODE == 2) keep = zv[j] > 0.f;
      if (!keep) g[j] = 0.f;
      const float xhat = (yv[j] - mu[j]) * is[j];
      o[j] = sc[j] * (g[j] - k1[j] - xhat * k2[j]);
    }
    store8(dy + at, o);
    if (dres) store8(dres + at, g);
  };
  long long r = (long long)blockIdx.x * l.rpi + lr;
  for (; r + (U - 1) * stride < M; r += U * stride) {
    float g[U][8], yv[U][8], zv[U][MODE == 2 ? 8 : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long at = (r + u * stride) * C + c8 * 8;
      load8(dz + at, g[u]);
      load8(y + at, yv[u]);
      if constexpr (MODE == 2) load8(z + at, zv[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one((r + u * stride) * C + c8 * 8, g[u], yv[u], zv[u]);
  }
  for (; r < M; r += stride) {
    const long long at = r * C + c8 * 8;
    float g[8], yv[8], zv[MODE == 2 ? 8 : 1];
    load8(dz + at, g);
    load8(y + at, yv);
    if constexpr (MODE == 2) load8(z + at, zv);
    one(at, g, yv, zv);
  }
}

static int bn_apply_grid(long long M, int C) {
  const BnLayout l = bn_layout(C);
  const long long need = ceil_div_ll(M, l.rpi);
  const long long cap = (long long)kNumSMs * 16;
  return (int)(need < cap ? need : cap);
}

template <typename T>
static int run_stats(const void* x, long long M, int C, long long ld, const float* gamma, const float* beta, float eps,
                     float momentum, float* rm, float* rv, float* stats, float* ws, cudaStream_t st) {
  StatsOp<T> op;
  op.x = (const T*)x;
  op.ld = ld;
  const int grid = bn_grid(M, C);
  bn_reduce_kernel<T, StatsOp<T>><<<grid, BN_THREADS, 0, st>>>(op, M, C, ws);
  int rc = check_launch("tdeed_bn_stats(partial)");
  if (rc) return rc;
  bn_stats_final_kernel<T><<<ceil_div(C, 8), 256, 0, st>>>((const T*)x, ws, grid, M, C, gamma, beta, eps, momentum, rm, rv, stats);
  return check_launch("tdeed_bn_stats(final)");
}

template <typename T, int MODE, int KB, int U>
static int run_bwd_cfg(const void* dz, const void* z, const void* y, long long M, int C, const float* stats, float* dgamma,
                        float* dbeta, void* dy, void* dres, float* ws, cudaStream_t st) {
  BwdOp<T, MODE, KB> op;
  op.dz = (const T*)dz;
  op.z = (const T*)z;
  op.y = (const T*)y;
  op.mean = stats;
  op.invstd = stats + C;
  op.C = C;
  const int grid = bn_grid(M, C);
  const int cpad = ((C + 7) / 8) * 8;
  float* coef = ws + (size_t)BN_MAX_GRID * 2 * cpad;
  bn_reduce_kernel<T, BwdOp<T, MODE, KB>><<<grid, BN_THREADS, 0, st>>>(op, M, C, ws);
  int rc = check_launch("tdeed_bn_act_bwd(partial)");
  if (rc) return rc;
  bn_bwd_final_kernel<<<ceil_div(C, 8), 256, 0, st>>>(ws, grid, M, C, dgamma, dbeta, coef);
  rc = check_launch("tdeed_bn_act_bwd(final)");
  if (rc) return rc;
  bn_act_bwd_apply_kernel<T, MODE, U><<<bn_apply_grid(M, C), BN_THREADS, 0, st>>>((const T*)dz, (const T*)z, (const T*)y, M, C, stats,
                                                                               coef, (T*)dy, (T*)dres);
  return check_launch("tdeed_bn_act_bwd(apply)");
}

template <typename T, int MODE>
static int run_bwd_mode(const void* dz, const void* z, const void* y, long long M, int C, const float* stats, float* dgamma,
                        float* dbeta, void* dy, void* dres, float* ws, cudaStream_t st) {
  // rows in flight per thread: reduction 4 (TDEED_BN_REDUCE_ROWS=8 to try 8: measured equal), apply 4 (TDEED_BN_APPLY_ROWS=2:
  // 11.66 vs 11.32 ms per FineGym_big step)
  static int kb = 0, u = 0;
  if (!kb) { const char* e = tdeed::dev_env("TDEED_BN_REDUCE_ROWS"); kb = (e && atoi(e) == 8) ? 8 : 4; }
  if (!u) { const char* e = tdeed::dev_env("TDEED_BN_APPLY_ROWS"); u = (e && atoi(e) == 2) ? 2 : 4; }
  if (kb == 8) return u == 4 ? run_bwd_cfg<T, MODE, 8, 4>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, ws, st)
                             : run_bwd_cfg<T, MODE, 8, 2>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, ws, st);
  return u == 4 ? run_bwd_cfg<T, MODE, 4, 4>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, ws, st)
                : run_bwd_cfg<T, MODE, 4, 2>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, ws, st);
}

template <typename T>
static int run_bwd(const void* dz, const void* z, const void* y, long long M, int C, const float* stats, float* dgamma,
                   float* dbeta, void* dy, void* dres, float* ws, cudaStream_t st) {
  if (!z) return run_bwd_mode<T, 0>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, ws, st);
  if (z == y) return run_bwd_mode<T, 1>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, ws, st);
  return run_bwd_mode<T, 2>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, ws, st);
}

}  // namespace tdeed

extern "C" long long tdeed_bn_workspace_floats(int C) {
  const long long cpad = ((C + 7) / 8) * 8;
  return (long long)tdeed::BN_MAX_GRID * 2 * cpad + 2 * cpad;
}

extern "C" int tdeed_bn_stats(int dtype, const void* x, long long M, int C, long long ld, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                              float* stats, float* workspace, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && stats && workspace, TDEED_ERR_SHAPE, "tdeed_bn_stats: null pointer");
  TDEED_REQUIRE(M > 0 && C > 0 && C <= 2048 && ld >= C && ld % 8 == 0 && (running_mean == nullptr) == (running_var == nullptr),
                TDEED_ERR_SHAPE, "tdeed_bn_stats: bad shape M=%lld C=%d ld=%lld", M, C, ld);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16) return run_stats<__nv_bfloat16>(x, M, C, ld, gamma, beta, eps, momentum, running_mean, running_var, stats, workspace, st);
  if (dtype == TDEED_F32) return run_stats<float>(x, M, C, ld, gamma, beta, eps, momentum, running_mean, running_var, stats, workspace, st);
  set_error("tdeed_bn_stats: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}

extern "C" int tdeed_bn_act_fwd(int dtype, const void* y, long long M, int C, const float* stats, const void* residual,
                                int relu, void* out, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(y && stats && out, TDEED_ERR_SHAPE, "tdeed_bn_act_fwd: null pointer");
  TDEED_REQUIRE(M > 0 && C > 0 && C % 8 == 0, TDEED_ERR_SHAPE, "tdeed_bn_act_fwd: bad shape M=%lld C=%d", M, C);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = bn_apply_grid(M, C);
  if (dtype == TDEED_BF16)
    bn_act_fwd_kernel<__nv_bfloat16><<<grid, BN_THREADS, 0, st>>>((const __nv_bfloat16*)y, M, C, stats + 2 * C, stats + 3 * C,
                                                                   (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)out);
  else if (dtype == TDEED_F32)
    bn_act_fwd_kernel<float><<<grid, BN_THREADS, 0, st>>>((const float*)y, M, C, stats + 2 * C, stats + 3 * C,
                                                           (const float*)residual, relu, (float*)out);
  else { set_error("tdeed_bn_act_fwd: dtype %d", dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_bn_act_fwd");
}

extern "C" int tdeed_bn_act_bwd(int dtype, const void* dz, const void* z, const void* y, long long M, int C,
                                const float* stats, float* dgamma, float* dbeta, void* dy, void* dres, float* workspace,
                                void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(dz && y && stats && dy && workspace, TDEED_ERR_SHAPE, "tdeed_bn_act_bwd: null pointer");
  TDEED_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 2048, TDEED_ERR_SHAPE, "tdeed_bn_act_bwd: bad shape M=%lld C=%d", M, C);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16) return run_bwd<__nv_bfloat16>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, workspace, st);
  if (dtype == TDEED_F32) return run_bwd<float>(dz, z, y, M, C, stats, dgamma, dbeta, dy, dres, workspace, st);
  set_error("tdeed_bn_act_bwd: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}
