// (7)(8)(9) SGP / SGP-Mixer token mixing and GroupNorm on [B, T, C] fp32 sequences.
// Reference: model/modules.py:159-186 (SGPBlock), :283-307 (SGPMixer), :348-363 (channel LayerNorm),
// :64,76 (AdaptiveMaxPool1d), :236,288 (linear upsample, align_corners=True).
//
// Grid = (16 GroupNorm groups, B clips).  A CTA keeps its group's [T, C/16] tile of the LayerNorm
// output in shared memory: depthwise temporal convolutions are sliding windows over that tile, the
// channel LayerNorm statistics (a reduction over ALL channels of a row) are computed per row by one
// warp with shuffles, and the GroupNorm statistics are CTA-local because a CTA owns a whole group.
// convw and convkw read the same input, so their kernels are merged into one `up`-tap kernel.
// These tensors are tiny (B*T*C <= a few MB): the kernels are latency-bound, the design goal is few launches.
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace tdeed {

constexpr int SG_THREADS = 256;
constexpr int SG_GROUPS = 16;
constexpr float SG_EPS = 1e-5f;

struct PoolWin { int s, e; };
__device__ inline PoolWin pool_window(int t, int t_in, int t_out) {   // AdaptiveMaxPool1d window
  PoolWin wdw;
  wdw.s = (int)(((long long)t * t_in) / t_out);
  wdw.e = (int)((((long long)(t + 1)) * t_in + t_out - 1) / t_out);
  return wdw;
}
__device__ inline float pooled(const float* __restrict__ xb, int C, int c, PoolWin wdw) {
  float m = xb[(size_t)wdw.s * C + c];
  for (int r = wdw.s + 1; r < wdw.e; ++r) m = fmaxf(m, xb[(size_t)r * C + c]);
  return m;
}

// LayerNorm statistics over channels for rows [0, t_out) of the (pooled) sequence; warp per row.
__device__ inline void row_stats(const float* __restrict__ xb, int C, int t_in, int t_out, float* s_mean, float* s_rstd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < t_out; t += SG_THREADS / 32) {
    const PoolWin wdw = pool_window(t, t_in, t_out);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += pooled(xb, C, c, wdw);
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float d = pooled(xb, C, c, wdw) - mean;
      q = fmaf(d, d, q);
    }
    const float var = warp_sum(q) / (float)C;
    if (lane == 0) {
      s_mean[t] = mean;
      s_rstd[t] = 1.f / sqrtf(var + SG_EPS);
    }
  }
}

// Cluster version: the 16 CTAs of one clip (one per GroupNorm group) form a thread-block cluster; each reduces ITS cg channels
// of every row and the partial sums are exchanged through distributed shared memory — the LayerNorm statistics (a reduction
// over all C channels) cost T*C/16 loads per CTA instead of T*C (the stand-alone version recomputes them in every CTA, which
// was 94 % of the kernel's memory traffic).  s_part: [T] floats of this CTA, readable by its cluster peers.
__device__ inline void row_stats_cluster(const float* __restrict__ xb, int C, int c0, int cg_, int t_in, int t_out, float* s_part,
                                         float* s_mean, float* s_rstd) {
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nrank = cluster.num_blocks();
  for (int t = warp; t < t_out; t += SG_THREADS / 32) {
    const PoolWin wdw = pool_window(t, t_in, t_out);
    float s = 0.f;
    for (int cl = lane; cl < cg_; cl += 32) s += pooled(xb, C, c0 + cl, wdw);
    s = warp_sum(s);
    if (lane == 0) s_part[t] = s;
  }
  cluster.sync();
  for (int t = threadIdx.x; t < t_out; t += SG_THREADS) {
    float m = 0.f;
    for (unsigned r = 0; r < nrank; ++r) m += cluster.map_shared_rank(s_part, r)[t];
    s_mean[t] = m / (float)C;
  }
  cluster.sync();                                  // all peers have read s_part; s_mean is complete
  for (int t = warp; t < t_out; t += SG_THREADS / 32) {
    const PoolWin wdw = pool_window(t, t_in, t_out);
    const float mean = s_mean[t];
    float q = 0.f;
    for (int cl = lane; cl < cg_; cl += 32) {
      const float d = pooled(xb, C, c0 + cl, wdw) - mean;
      q = fmaf(d, d, q);
    }
    q = warp_sum(q);
    if (lane == 0) s_part[t] = q;
  }
  cluster.sync();
  for (int t = threadIdx.x; t < t_out; t += SG_THREADS) {
    float v = 0.f;
    for (unsigned r = 0; r < nrank; ++r) v += cluster.map_shared_rank(s_part, r)[t];
    s_rstd[t] = 1.f / sqrtf(v / (float)C + SG_EPS);
  }
  cluster.sync();                                  // nobody may leave (or reuse s_part) while peers still read it
}

// depthwise conv over time on a [T][cg] smem tile, zero padded; weights [cg][k] in smem
__device__ inline float dwconv(const float* __restrict__ tile, int T, int cg, int t, int cl,
                               const float* __restrict__ wk, int k, float bias) {
  const int hk = k / 2;
  float a = bias;
  const int lo = max(0, t - hk), hi = min(T - 1, t + hk);
  const float* wrow = wk + cl * k + (lo - t + hk);
  const float* src = tile + lo * cg + cl;
  for (int r = lo; r <= hi; ++r, ++wrow, src += cg) a = fmaf(*wrow, *src, a);
  return a;
}

struct SgpW {
  tdeed_sgp_weights w;
};

template <bool CLUSTER>
__global__ void __launch_bounds__(SG_THREADS)
sgp_mix_kernel(const float* __restrict__ x, int t_in, int T, int C, int ks, int up, SgpW W,
               float* __restrict__ y, void* __restrict__ g, int g_dtype, int big) {
  extern __shared__ float smem[];
  const int cg = C / SG_GROUPS;
  const int grp = blockIdx.x, b = blockIdx.y;
  const int c0 = grp * cg;
  float* s_mean = smem;                 // [T]
  float* s_rstd = s_mean + T;           // [T]
  float* s_ln = s_rstd + T;             // [T][cg]
  // big (long sequences, T*C/16 tiles that do not fit twice): no s_x tile — the pooled input is recomputed and y is re-read
  // from global memory (L2) for the GroupNorm passes
  float* s_x = s_ln + T * cg;           // [T][cg]  pooled input, later y
  float* s_psi = big ? s_x : s_x + T * cg;          // [cg][ks]
  float* s_mrg = s_psi + cg * ks;       // [cg][up]  convkw with convw folded into the centre taps
  float* s_phi = s_mrg + cg * up;       // [cg]
  float* s_red = s_phi + cg;            // [32]
  float* s_part = s_red + 32;           // [T]  (cluster version only)
  const float* xb = x + (size_t)b * t_in * C;
  const tdeed_sgp_weights& w = W.w;

  if (CLUSTER) row_stats_cluster(xb, C, c0, cg, t_in, T, s_part, s_mean, s_rstd);
  else row_stats(xb, C, t_in, T, s_mean, s_rstd);
  for (int i = threadIdx.x; i < cg * ks; i += SG_THREADS) s_psi[i] = w.psi_w[(size_t)c0 * ks + i];
  for (int i = threadIdx.x; i < cg * up; i += SG_THREADS) {
    const int cl = i / up, k = i - cl * up;
    float v = w.convkw_w[(size_t)(c0 + cl) * up + k];
    const int kk = k - (up / 2 - ks / 2);
    if (kk >= 0 && kk < ks) v += w.convw_w[(size_t)(c0 + cl) * ks + kk];
    s_mrg[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg;
    const float xv = pooled(xb, C, c0 + cl, pool_window(t, t_in, T));
    if (!big) s_x[i] = xv;
    s_ln[i] = (xv - s_mean[t]) * s_rstd[t] * w.ln_w[c0 + cl] + w.ln_b[c0 + cl];
  }
  __syncthreads();
  for (int cl = threadIdx.x; cl < cg; cl += SG_THREADS) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += s_ln[t * cg + cl];
    s_phi[cl] = fmaxf(fmaf(w.gfc_w[c0 + cl], s / (float)T, w.gfc_b[c0 + cl]), 0.f);
  }
  __syncthreads();
  float lsum = 0.f;
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg, c = c0 + cl;
    const float ln = s_ln[i];
    const float psi = dwconv(s_ln, T, cg, t, cl, s_psi, ks, w.psi_b[c]);
    const float win = dwconv(s_ln, T, cg, t, cl, s_mrg, up, w.convw_b[c] + w.convkw_b[c]);
    const float fc = fmaf(w.fc_w[c], ln, w.fc_b[c]);
    const float xin = big ? pooled(xb, C, c, pool_window(t, t_in, T)) : s_x[i];
    const float yv = xin + (fc * s_phi[cl] + win * psi + ln);
    if (!big) s_x[i] = yv;
    y[((size_t)b * T + t) * C + c] = yv;
    lsum += yv;
  }
  const float n = (float)(T * cg);
  const float mean = block_sum(lsum, s_red) / n;
  float lq = 0.f;
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {   // (block_sum's barriers made this CTA's y stores visible)
    const float d = (big ? y[((size_t)b * T + i / cg) * C + c0 + i % cg] : s_x[i]) - mean;
    lq = fmaf(d, d, lq);
  }
  const float rstd = 1.f / sqrtf(block_sum(lq, s_red) / n + SG_EPS);
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg, c = c0 + cl;
    const size_t o = ((size_t)b * T + t) * C + c;
    const float gv = ((big ? y[o] : s_x[i]) - mean) * rstd * w.gn_w[c] + w.gn_b[c];
    if (g_dtype == TDEED_F32) reinterpret_cast<float*>(g)[o] = gv;
    else reinterpret_cast<__nv_bfloat16*>(g)[o] = __float2bfloat16_rn(gv);
  }
}

struct MixW {
  tdeed_mixer_weights w;
};

__device__ inline void st_cat(void* cat, int dtype, size_t o, float v) {
  if (dtype == TDEED_F32) reinterpret_cast<float*>(cat)[o] = v;
  else reinterpret_cast<__nv_bfloat16*>(cat)[o] = __float2bfloat16_rn(v);
}

template <bool CLUSTER>
__global__ void __launch_bounds__(SG_THREADS)
sgp_mixer_kernel(const float* __restrict__ xc, const float* __restrict__ skip, int tc, int T, int C, int ks, int up,
                 MixW W, void* __restrict__ cat, int cat_dtype, int big) {
  extern __shared__ float smem[];
  const int cg = C / SG_GROUPS;
  const int grp = blockIdx.x, b = blockIdx.y;
  const int c0 = grp * cg;
  float* s_mz = smem;                  // [T] mean / rstd of skip rows
  float* s_rz = s_mz + T;
  float* s_mx = s_rz + T;              // [tc] mean / rstd of coarse rows
  float* s_rx = s_mx + tc;
  float* s_z = s_rx + tc;              // [T][cg]  LN1(skip)
  // big (long sequences): ONE [T][cg] tile, used for z (outputs 1, 3, 5) and then re-filled with u (outputs 2, 4, 6); the
  // upsampled rows are interpolated straight from global memory, so the coarse tile s_c is not needed either
  float* s_u = big ? s_z : s_z + T * cg;           // [T][cg]  upsampled LN2(x)
  float* s_c = s_u + T * cg;           // [tc][cg] LN2(x)
  float* s_psi1 = big ? s_c : s_c + tc * cg;       // [cg][ks]
  float* s_psi2 = s_psi1 + cg * ks;
  float* s_m1 = s_psi2 + cg * ks;      // [cg][up]
  float* s_m2 = s_m1 + cg * up;
  float* s_phi1 = s_m2 + cg * up;      // [cg]
  float* s_phi2 = s_phi1 + cg;
  float* s_part = s_phi2 + cg;         // [T]  (cluster version only)
  const tdeed_mixer_weights& w = W.w;
  const float* zb = skip + (size_t)b * T * C;
  const float* xb = xc + (size_t)b * tc * C;

  if (CLUSTER) {
    row_stats_cluster(zb, C, c0, cg, T, T, s_part, s_mz, s_rz);
    row_stats_cluster(xb, C, c0, cg, tc, tc, s_part, s_mx, s_rx);
  } else {
    row_stats(zb, C, T, T, s_mz, s_rz);
    row_stats(xb, C, tc, tc, s_mx, s_rx);
  }
  for (int i = threadIdx.x; i < cg * ks; i += SG_THREADS) {
    s_psi1[i] = w.psi1_w[(size_t)c0 * ks + i];
    s_psi2[i] = w.psi2_w[(size_t)c0 * ks + i];
  }
  for (int i = threadIdx.x; i < cg * up; i += SG_THREADS) {
    const int cl = i / up, k = i - cl * up;
    float v1 = w.convkw1_w[(size_t)(c0 + cl) * up + k], v2 = w.convkw2_w[(size_t)(c0 + cl) * up + k];
    const int kk = k - (up / 2 - ks / 2);
    if (kk >= 0 && kk < ks) {
      v1 += w.convw1_w[(size_t)(c0 + cl) * ks + kk];
      v2 += w.convw2_w[(size_t)(c0 + cl) * ks + kk];
    }
    s_m1[i] = v1;
    s_m2[i] = v2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg, c = c0 + cl;
    s_z[i] = (zb[(size_t)t * C + c] - s_mz[t]) * s_rz[t] * w.ln1_w[c] + w.ln1_b[c];
  }
  const float scale = (T > 1) ? (float)(tc - 1) / (float)(T - 1) : 0.f;
  const size_t ldc = (size_t)6 * C;
  if (big) {
    __syncthreads();
    // ---- pass 1: z tile ----
    for (int cl = threadIdx.x; cl < cg; cl += SG_THREADS) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += s_z[t * cg + cl];
      s_phi1[cl] = fmaxf(fmaf(w.gfc1_w[c0 + cl], s / (float)T, w.gfc1_b[c0 + cl]), 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
      const int cl = i % cg, t = i / cg, c = c0 + cl;
      const float zv = s_z[i];
      const float o1 = dwconv(s_z, T, cg, t, cl, s_m1, up, w.convw1_b[c] + w.convkw1_b[c]) *
                       dwconv(s_z, T, cg, t, cl, s_psi1, ks, w.psi1_b[c]);
      const size_t row = ((size_t)b * T + t) * ldc + c;
      st_cat(cat, cat_dtype, row, o1);
      st_cat(cat, cat_dtype, row + 2 * (size_t)C, fmaf(w.fc1_w[c], zv, w.fc1_b[c]) * s_phi1[cl]);
      st_cat(cat, cat_dtype, row + 4 * (size_t)C, zv);
    }
    __syncthreads();
    // ---- pass 2: u tile in the same shared memory ----
    for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
      const int cl = i % cg, t = i / cg, c = c0 + cl;
      const float real = scale * (float)t;
      const int i0 = (int)real;
      const int i1 = i0 + ((i0 < tc - 1) ? 1 : 0);
      const float l1 = fminf(fmaxf(real - (float)i0, 0.f), 1.f), l0 = 1.f - l1;
      const float a0 = (xb[(size_t)i0 * C + c] - s_mx[i0]) * s_rx[i0] * w.ln2_w[c] + w.ln2_b[c];
      const float a1 = (xb[(size_t)i1 * C + c] - s_mx[i1]) * s_rx[i1] * w.ln2_w[c] + w.ln2_b[c];
      s_u[i] = l0 * a0 + l1 * a1;
    }
    __syncthreads();
    for (int cl = threadIdx.x; cl < cg; cl += SG_THREADS) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s += s_u[t * cg + cl];
      s_phi2[cl] = fmaxf(fmaf(w.gfc2_w[c0 + cl], s / (float)T, w.gfc2_b[c0 + cl]), 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
      const int cl = i % cg, t = i / cg, c = c0 + cl;
      const float uv = s_u[i];
      const float o2 = dwconv(s_u, T, cg, t, cl, s_m2, up, w.convw2_b[c] + w.convkw2_b[c]) *
                       dwconv(s_u, T, cg, t, cl, s_psi2, ks, w.psi2_b[c]);
      const size_t row = ((size_t)b * T + t) * ldc + c;
      st_cat(cat, cat_dtype, row + C, o2);
      st_cat(cat, cat_dtype, row + 3 * (size_t)C, fmaf(w.fc2_w[c], uv, w.fc2_b[c]) * s_phi2[cl]);
      st_cat(cat, cat_dtype, row + 5 * (size_t)C, uv);
    }
    return;
  }
  for (int i = threadIdx.x; i < tc * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg, c = c0 + cl;
    s_c[i] = (xb[(size_t)t * C + c] - s_mx[t]) * s_rx[t] * w.ln2_w[c] + w.ln2_b[c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg;
    const float real = scale * (float)t;
    const int i0 = (int)real;
    const int i1 = i0 + ((i0 < tc - 1) ? 1 : 0);
    const float l1 = fminf(fmaxf(real - (float)i0, 0.f), 1.f), l0 = 1.f - l1;
    s_u[i] = l0 * s_c[i0 * cg + cl] + l1 * s_c[i1 * cg + cl];
  }
  __syncthreads();
  for (int q = threadIdx.x; q < 2 * cg; q += SG_THREADS) {
    const int which = q / cg, cl = q - which * cg, c = c0 + cl;
    const float* tile = which ? s_u : s_z;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += tile[t * cg + cl];
    const float m = s / (float)T;
    if (which) s_phi2[cl] = fmaxf(fmaf(w.gfc2_w[c], m, w.gfc2_b[c]), 0.f);
    else s_phi1[cl] = fmaxf(fmaf(w.gfc1_w[c], m, w.gfc1_b[c]), 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg, c = c0 + cl;
    const float zv = s_z[i], uv = s_u[i];
    const float o1 = dwconv(s_z, T, cg, t, cl, s_m1, up, w.convw1_b[c] + w.convkw1_b[c]) *
                     dwconv(s_z, T, cg, t, cl, s_psi1, ks, w.psi1_b[c]);
    const float o2 = dwconv(s_u, T, cg, t, cl, s_m2, up, w.convw2_b[c] + w.convkw2_b[c]) *
                     dwconv(s_u, T, cg, t, cl, s_psi2, ks, w.psi2_b[c]);
    const float o3 = fmaf(w.fc1_w[c], zv, w.fc1_b[c]) * s_phi1[cl];
    const float o4 = fmaf(w.fc2_w[c], uv, w.fc2_b[c]) * s_phi2[cl];
    const size_t row = ((size_t)b * T + t) * ldc + c;
    st_cat(cat, cat_dtype, row, o1);
    st_cat(cat, cat_dtype, row + C, o2);
    st_cat(cat, cat_dtype, row + 2 * (size_t)C, o3);
    st_cat(cat, cat_dtype, row + 3 * (size_t)C, o4);
    st_cat(cat, cat_dtype, row + 4 * (size_t)C, zv);
    st_cat(cat, cat_dtype, row + 5 * (size_t)C, uv);
  }
}

__global__ void __launch_bounds__(SG_THREADS)
groupnorm_kernel(const float* __restrict__ x, int T, int C, int groups, const float* __restrict__ gamma,
                 const float* __restrict__ beta, void* __restrict__ out, int out_dtype) {
  extern __shared__ float smem[];
  const int cg = C / groups;
  const int c0 = blockIdx.x * cg, b = blockIdx.y;
  float* s_x = smem;            // [T][cg]
  float* s_red = s_x + T * cg;  // [32]
  float lsum = 0.f;
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg;
    const float v = x[((size_t)b * T + t) * C + c0 + cl];
    s_x[i] = v;
    lsum += v;
  }
  const float n = (float)(T * cg);
  const float mean = block_sum(lsum, s_red) / n;
  float lq = 0.f;
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const float d = s_x[i] - mean;
    lq = fmaf(d, d, lq);
  }
  const float rstd = 1.f / sqrtf(block_sum(lq, s_red) / n + SG_EPS);
  for (int i = threadIdx.x; i < T * cg; i += SG_THREADS) {
    const int cl = i % cg, t = i / cg, c = c0 + cl;
    st_cat(out, out_dtype, ((size_t)b * T + t) * C + c, (s_x[i] - mean) * rstd * gamma[c] + beta[c]);
  }
}

static int set_smem(const void* fn, size_t smem, const char* what, size_t* cur) {
  if (smem > *cur) {
    TDEED_REQUIRE(smem <= 227 * 1024, TDEED_ERR_UNSUPPORTED, "%s: needs %zu B of shared memory (T*C/16 too large)", what, smem);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    *cur = smem;
  }
  return TDEED_OK;
}

// Launch `kern` as clusters of SG_GROUPS (= 16, a non-portable size) CTAs along x.  Returns false when this device / shared
// memory size cannot co-schedule such a cluster (the caller then uses the stand-alone kernel).
template <typename... Args>
static bool launch_cluster16(void (*kern)(Args...), dim3 grid, size_t smem, cudaStream_t st, Args... args) {
  static_assert(SG_GROUPS == 16, "cluster size");
  // per kernel instantiation: the attribute is set once, the co-scheduling query is cached for the largest smem size seen
  static int allowed = -1;
  static size_t ok_smem = 0, bad_smem = ~(size_t)0;
  if (allowed < 0) {
    allowed = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? 1 : 0;
    if (!allowed) cudaGetLastError();
  }
  if (!allowed || smem >= bad_smem) return false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(SG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = SG_GROUPS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (smem > ok_smem) {
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, (const void*)kern, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      bad_smem = smem;
      return false;
    }
    ok_smem = smem;
  }
  if (cudaLaunchKernelEx(&cfg, kern, args...) != cudaSuccess) {
    cudaGetLastError();
    bad_smem = smem;
    return false;
  }
  return true;
}

static bool sgp_no_cluster() {
  static int v = -1;
  if (v < 0) {
    const char* e = tdeed::dev_env("TDEED_SGP_NO_CLUSTER");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace tdeed

extern "C" int tdeed_sgp_mix_fwd(const float* x, int B, int t_in, int t_out, int C, int ks, int up,
                                 const tdeed_sgp_weights* w_host, float* y, void* g, int g_dtype, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && w_host && y && g, TDEED_ERR_SHAPE, "tdeed_sgp_mix_fwd: null pointer");
  TDEED_REQUIRE(B > 0 && B <= 65535 && t_out > 0 && t_in >= t_out && C % SG_GROUPS == 0 && ks % 2 == 1 && up % 2 == 1 && up >= ks,
                TDEED_ERR_SHAPE, "tdeed_sgp_mix_fwd: bad shape B=%d t_in=%d t_out=%d C=%d ks=%d up=%d", B, t_in, t_out, C, ks, up);
  const int cg = C / SG_GROUPS;
  size_t smem = ((size_t)2 * t_out + 2 * (size_t)t_out * cg + (size_t)cg * (ks + up + 1) + 32) * sizeof(float);
  const int big = smem > 200 * 1024;
  if (big) smem -= (size_t)t_out * cg * sizeof(float);
  smem += (size_t)t_out * sizeof(float);           // s_part of the cluster version
  static size_t cur = 48 * 1024, cur_cl = 48 * 1024;
  SgpW W{*w_host};
  if (!sgp_no_cluster() && set_smem((const void*)sgp_mix_kernel<true>, smem, "tdeed_sgp_mix_fwd", &cur_cl) == TDEED_OK &&
      launch_cluster16(sgp_mix_kernel<true>, dim3(SG_GROUPS, B), smem, (cudaStream_t)stream, x, t_in, t_out, C, ks, up, W, y, g, g_dtype, big))
    return check_launch("tdeed_sgp_mix_fwd(cluster)");
  int rc = set_smem((const void*)sgp_mix_kernel<false>, smem, "tdeed_sgp_mix_fwd", &cur);
  if (rc) return rc;
  sgp_mix_kernel<false><<<dim3(SG_GROUPS, B), SG_THREADS, smem, (cudaStream_t)stream>>>(x, t_in, t_out, C, ks, up, W, y, g, g_dtype, big);
  return check_launch("tdeed_sgp_mix_fwd");
}

extern "C" int tdeed_sgp_mixer_mix_fwd(const float* x_coarse, const float* skip, int B, int t_coarse, int T, int C,
                                       int ks, int up, const tdeed_mixer_weights* w_host, void* cat, int cat_dtype,
                                       void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x_coarse && skip && w_host && cat, TDEED_ERR_SHAPE, "tdeed_sgp_mixer_mix_fwd: null pointer");
  TDEED_REQUIRE(B > 0 && B <= 65535 && T > 0 && t_coarse > 0 && C % SG_GROUPS == 0 && ks % 2 == 1 && up % 2 == 1 && up >= ks,
                TDEED_ERR_SHAPE, "tdeed_sgp_mixer_mix_fwd: bad shape B=%d tc=%d T=%d C=%d", B, t_coarse, T, C);
  const int cg = C / SG_GROUPS;
  size_t smem = ((size_t)2 * T + 2 * t_coarse + (size_t)(2 * T + t_coarse) * cg + (size_t)cg * (2 * ks + 2 * up + 2)) * sizeof(float);
  const int big = smem > 200 * 1024;
  if (big) smem -= (size_t)(T + t_coarse) * cg * sizeof(float);
  smem += (size_t)T * sizeof(float);               // s_part of the cluster version
  static size_t cur = 48 * 1024, cur_cl = 48 * 1024;
  MixW W{*w_host};
  if (!sgp_no_cluster() && set_smem((const void*)sgp_mixer_kernel<true>, smem, "tdeed_sgp_mixer_mix_fwd", &cur_cl) == TDEED_OK &&
      launch_cluster16(sgp_mixer_kernel<true>, dim3(SG_GROUPS, B), smem, (cudaStream_t)stream, x_coarse, skip, t_coarse, T, C, ks, up, W,
                       cat, cat_dtype, big))
    return check_launch("tdeed_sgp_mixer_mix_fwd(cluster)");
  int rc = set_smem((const void*)sgp_mixer_kernel<false>, smem, "tdeed_sgp_mixer_mix_fwd", &cur);
  if (rc) return rc;
  sgp_mixer_kernel<false><<<dim3(SG_GROUPS, B), SG_THREADS, smem, (cudaStream_t)stream>>>(x_coarse, skip, t_coarse, T, C, ks, up, W, cat, cat_dtype, big);
  return check_launch("tdeed_sgp_mixer_mix_fwd");
}

extern "C" int tdeed_groupnorm_fwd(const float* x, int B, int T, int C, int groups, const float* gamma, const float* beta,
                                   void* out, int out_dtype, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && gamma && beta && out, TDEED_ERR_SHAPE, "tdeed_groupnorm_fwd: null pointer");
  TDEED_REQUIRE(B > 0 && B <= 65535 && T > 0 && groups > 0 && C % groups == 0, TDEED_ERR_SHAPE,
                "tdeed_groupnorm_fwd: bad shape B=%d T=%d C=%d groups=%d", B, T, C, groups);
  const size_t smem = ((size_t)T * (C / groups) + 32) * sizeof(float);
  static size_t cur = 48 * 1024;
  int rc = set_smem((const void*)groupnorm_kernel, smem, "tdeed_groupnorm_fwd", &cur);
  if (rc) return rc;
  groupnorm_kernel<<<dim3(groups, B), SG_THREADS, smem, (cudaStream_t)stream>>>(x, T, C, groups, gamma, beta, out, out_dtype);
  return check_launch("tdeed_groupnorm_fwd");
}
