// Backward of the ED-SGP-Mixer temporal layers (model/modules.py:58-318) on [B, T, C] fp32 sequences.
// The forward of the training step reuses the fused inference kernels (sgp.cu) and saves only the block inputs /
// residual streams; every backward kernel recomputes what it needs from those (the tensors are <= a few MB, the
// kernels are latency bound, so recomputation is cheaper than saving a dozen intermediates per block).
//
//   tdeed_chan_ln_fwd      (optional AdaptiveMaxPool1d) + channel LayerNorm  -> ln, pooled x, argmax rows
//   tdeed_chan_ln_bwd      LayerNorm backward (+ dgamma/dbeta) (+ scatter through the max-pool)
//   tdeed_sgp_branch_bwd   backward of  fc(ln)*relu(gfc(mean_T ln)) + (convw(ln)+convkw(ln))*psi(ln) [+ ln]
//   tdeed_groupnorm_bwd    GroupNorm(16) backward (+ dgamma/dbeta), fused with the residual add
//   tdeed_gelu_fwd / _bwd  exact-erf GELU on the MLP hidden activations
//   tdeed_upsample_bwd     backward of the align_corners linear upsample of the mixer
// All reductions run in a fixed order (deterministic); no atomics.
#include "train_reduce.cuh"

namespace tdeed {

constexpr int TS_THREADS = 256;
constexpr float TS_EPS = 1e-5f;

__device__ inline void pool_win(int t, int t_in, int t_out, int& s, int& e) {
  s = (int)(((long long)t * t_in) / t_out);
  e = (int)((((long long)(t + 1)) * t_in + t_out - 1) / t_out);
}

// ---- (pool +) LayerNorm forward.  warp per output row ----
__global__ void __launch_bounds__(TS_THREADS)
chan_ln_fwd_kernel(const float* __restrict__ x, int B, int t_in, int T, int C, const float* __restrict__ w,
                   const float* __restrict__ bia, float* __restrict__ xp, int* __restrict__ arg, float* __restrict__ ln,
                   float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (TS_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= B * T) return;
  const int b = row / T, t = row - b * T;
  int s, e;
  pool_win(t, t_in, T, s, e);
  const float* xb = x + (size_t)b * t_in * C;
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) {
    float m = xb[(size_t)s * C + c];
    int am = s;
    for (int r = s + 1; r < e; ++r) {
      const float v = xb[(size_t)r * C + c];
      if (v > m) { m = v; am = r; }          // first maximum wins, like torch's max-pool backward
    }
    if (xp) xp[(size_t)row * C + c] = m;
    if (arg) arg[(size_t)row * C + c] = am;
    sum += m;
  }
  const float mean = warp_sum(sum) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    float m = xb[(size_t)s * C + c];
    for (int r = s + 1; r < e; ++r) m = fmaxf(m, xb[(size_t)r * C + c]);
    const float d = m - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) / (float)C + TS_EPS);
  for (int c = lane; c < C; c += 32) {
    float m = xb[(size_t)s * C + c];
    for (int r = s + 1; r < e; ++r) m = fmaxf(m, xb[(size_t)r * C + c]);
    ln[(size_t)row * C + c] = (m - mean) * rstd * w[c] + bia[c];
  }
  if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
}

// ---- LayerNorm backward.  xp: the (pooled) LN input [B*T, C]; dln: gradient w.r.t. the LN output with leading dim ld ----
__global__ void __launch_bounds__(TS_THREADS)
chan_ln_bwd_kernel(const float* __restrict__ xp, const float* __restrict__ stats, const float* __restrict__ dln, long long ld,
                   int rows, int C, const float* __restrict__ w, const float* __restrict__ add, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (TS_THREADS / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float dh = dln[(size_t)row * ld + c] * w[c];
    const float xh = (xp[(size_t)row * C + c] - mean) * rstd;
    s1 += dh;
    s2 = fmaf(dh, xh, s2);
  }
  s1 = warp_sum(s1) / (float)C;
  s2 = warp_sum(s2) / (float)C;
  for (int c = lane; c < C; c += 32) {
    const float dh = dln[(size_t)row * ld + c] * w[c];
    const float xh = (xp[(size_t)row * C + c] - mean) * rstd;
    float v = rstd * (dh - s1 - xh * s2);
    if (add) v += add[(size_t)row * C + c];
    dx[(size_t)row * C + c] = v;
  }
}

// dgamma[c] = sum_rows dln * xhat, dbeta[c] = sum_rows dln.  thread per channel, serial over rows (<= a few thousand)
__global__ void chan_ln_param_kernel(const float* __restrict__ xp, const float* __restrict__ stats, const float* __restrict__ dln,
                                     long long ld, int rows, int C, float* __restrict__ dw, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float d = dln[(size_t)r * ld + c];
    a = fmaf(d, (xp[(size_t)r * C + c] - stats[2 * r]) * stats[2 * r + 1], a);
    b += d;
  }
  dw[c] = a;
  db[c] = b;
}

// gradient through AdaptiveMaxPool1d: dx[b, r, c] = sum over the output rows t whose window holds r and whose argmax is r
__global__ void maxpool_bwd_kernel(const float* __restrict__ dxp, const int* __restrict__ arg, int B, int t_in, int T, int C,
                                   const float* __restrict__ add, float* __restrict__ dx) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * t_in * C) return;
  const int c = (int)(idx % C);
  const int r = (int)((idx / C) % t_in);
  const int b = (int)(idx / ((long long)C * t_in));
  float v = add ? add[idx] : 0.f;
  // candidate output rows: t with s(t) <= r < e(t);  t is within +-1 of r*T/t_in
  const int tc = (int)(((long long)r * T) / t_in);
  for (int t = max(0, tc - 1); t <= min(T - 1, tc + 1); ++t) {
    int s, e;
    pool_win(t, t_in, T, s, e);
    if (r >= s && r < e && arg[((size_t)b * T + t) * C + c] == r) v += dxp[((size_t)b * T + t) * C + c];
  }
  dx[idx] = v;
}

// ---- branch backward ----
struct BranchArgs {
  const float* ln; long long ld_ln;       // branch input [B*T, C] (leading dim ld_ln)
  const float* d_conv; const float* d_fc; const float* d_id; long long ld_g;   // upstream gradients (d_id may be null)
  const float *psi_w, *psi_b, *convw_w, *convw_b, *convkw_w, *convkw_b, *fc_w, *fc_b, *gfc_w, *gfc_b;
  int B, T, C, ks, up;
  float *m, *dpre, *phi;                  // [B, C] scratch
  float *dcwk, *dpsi;                     // [B, T, C] scratch
  float* d_ln;                            // out [B, T, C]
  float *g_psi_w, *g_psi_b, *g_convw_w, *g_convw_b, *g_convkw_w, *g_convkw_b, *g_fc_w, *g_fc_b, *g_gfc_w, *g_gfc_b;
};

__device__ inline float dw_at(const float* __restrict__ src, long long ld, int T, int t, const float* __restrict__ wk, int k, float bias) {
  // depthwise conv over time at row t for one channel: src points at (b, 0, c)
  const int hk = k / 2;
  float a = bias;
  const int lo = max(0, t - hk), hi = min(T - 1, t + hk);
  for (int r = lo; r <= hi; ++r) a = fmaf(wk[r - t + hk], src[(size_t)r * ld], a);
  return a;
}

// k1: thread per (b, c): phi, dphi -> dpre
__global__ void branch_k1(BranchArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.C) return;
  const int c = i % a.C, b = i / a.C;
  const float* ln = a.ln + (size_t)b * a.T * a.ld_ln + c;
  const float* dfc = a.d_fc + (size_t)b * a.T * a.ld_g + c;
  float s = 0.f, dphi = 0.f;
  for (int t = 0; t < a.T; ++t) {
    const float l = ln[(size_t)t * a.ld_ln];
    s += l;
    dphi = fmaf(dfc[(size_t)t * a.ld_g], fmaf(a.fc_w[c], l, a.fc_b[c]), dphi);
  }
  const float m = s / (float)a.T;
  const float pre = fmaf(a.gfc_w[c], m, a.gfc_b[c]);
  a.m[i] = m;
  a.phi[i] = fmaxf(pre, 0.f);
  a.dpre[i] = pre > 0.f ? dphi : 0.f;
}

// k2: thread per (b, t, c): dcwk = d_conv * psi, dpsi = d_conv * (convw + convkw)
__global__ void branch_k2(BranchArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)a.B * a.T * a.C) return;
  const int c = (int)(i % a.C);
  const int t = (int)((i / a.C) % a.T);
  const int b = (int)(i / ((long long)a.C * a.T));
  const float* ln = a.ln + (size_t)b * a.T * a.ld_ln + c;
  const float psi = dw_at(ln, a.ld_ln, a.T, t, a.psi_w + (size_t)c * a.ks, a.ks, a.psi_b[c]);
  const float cwk = dw_at(ln, a.ld_ln, a.T, t, a.convw_w + (size_t)c * a.ks, a.ks, a.convw_b[c]) +
                    dw_at(ln, a.ld_ln, a.T, t, a.convkw_w + (size_t)c * a.up, a.up, a.convkw_b[c]);
  const float d = a.d_conv[((size_t)b * a.T + t) * a.ld_g + c];
  a.dcwk[i] = d * psi;
  a.dpsi[i] = d * cwk;
}

__device__ inline float dwT_at(const float* __restrict__ g, int C, int T, int t, const float* __restrict__ wk, int k) {
  // transposed depthwise conv: d_in[t] = sum_k w[k] * d_out[t - k + hk];  g points at (b, 0, c) of a [B,T,C] tensor
  const int hk = k / 2;
  float a = 0.f;
  for (int kk = 0; kk < k; ++kk) {
    const int r = t - kk + hk;
    if (r >= 0 && r < T) a = fmaf(wk[kk], g[(size_t)r * C], a);
  }
  return a;
}

// k3: thread per (b, t, c): d_ln
__global__ void branch_k3(BranchArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)a.B * a.T * a.C) return;
  const int c = (int)(i % a.C);
  const int t = (int)((i / a.C) % a.T);
  const int b = (int)(i / ((long long)a.C * a.T));
  const size_t bc = (size_t)b * a.C + c;
  const size_t base = (size_t)b * a.T * a.C + c;
  float v = a.d_id ? a.d_id[((size_t)b * a.T + t) * a.ld_g + c] : 0.f;
  v = fmaf(a.fc_w[c] * a.phi[bc], a.d_fc[((size_t)b * a.T + t) * a.ld_g + c], v);
  v += dwT_at(a.dpsi + base, a.C, a.T, t, a.psi_w + (size_t)c * a.ks, a.ks);
  v += dwT_at(a.dcwk + base, a.C, a.T, t, a.convw_w + (size_t)c * a.ks, a.ks);
  v += dwT_at(a.dcwk + base, a.C, a.T, t, a.convkw_w + (size_t)c * a.up, a.up);
  v += a.dpre[bc] * a.gfc_w[c] / (float)a.T;
  a.d_ln[i] = v;
}

// k4: thread per (c, slot): parameter gradients.  slots: [0,ks) psi_w, [ks,2ks) convw_w, [2ks,2ks+up) convkw_w, then
// psi_b, convw_b, convkw_b, fc_w, fc_b, gfc_w, gfc_b
__global__ void branch_k4(BranchArgs a) {
  const int nslot = 2 * a.ks + a.up + 7;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)a.C * nslot) return;
  const int c = (int)(i % a.C), slot = (int)(i / a.C);
  float s = 0.f;
  if (slot < 2 * a.ks + a.up) {
    const float* g;
    int k, kk;
    float* out;
    if (slot < a.ks) { g = a.dpsi; k = a.ks; kk = slot; out = a.g_psi_w + (size_t)c * a.ks + kk; }
    else if (slot < 2 * a.ks) { g = a.dcwk; k = a.ks; kk = slot - a.ks; out = a.g_convw_w + (size_t)c * a.ks + kk; }
    else { g = a.dcwk; k = a.up; kk = slot - 2 * a.ks; out = a.g_convkw_w + (size_t)c * a.up + kk; }
    const int off = kk - k / 2;
    for (int b = 0; b < a.B; ++b) {
      const float* ln = a.ln + (size_t)b * a.T * a.ld_ln + c;
      const float* gb = g + (size_t)b * a.T * a.C + c;
      const int lo = max(0, -off), hi = min(a.T - 1, a.T - 1 - off);
      for (int t = lo; t <= hi; ++t) s = fmaf(gb[(size_t)t * a.C], ln[(size_t)(t + off) * a.ld_ln], s);
    }
    *out = s;
    return;
  }
  const int which = slot - (2 * a.ks + a.up);
  if (which <= 2) {                      // biases of psi / convw / convkw
    const float* g = which == 0 ? a.dpsi : a.dcwk;
    for (int b = 0; b < a.B; ++b)
      for (int t = 0; t < a.T; ++t) s += g[((size_t)b * a.T + t) * a.C + c];
    (which == 0 ? a.g_psi_b : which == 1 ? a.g_convw_b : a.g_convkw_b)[c] = s;
  } else if (which <= 4) {               // fc_w, fc_b
    for (int b = 0; b < a.B; ++b) {
      const float phi = a.phi[(size_t)b * a.C + c];
      for (int t = 0; t < a.T; ++t) {
        const float d = a.d_fc[((size_t)b * a.T + t) * a.ld_g + c] * phi;
        s = which == 3 ? fmaf(d, a.ln[((size_t)b * a.T + t) * a.ld_ln + c], s) : s + d;
      }
    }
    (which == 3 ? a.g_fc_w : a.g_fc_b)[c] = s;
  } else {                               // gfc_w, gfc_b
    for (int b = 0; b < a.B; ++b) {
      const float d = a.dpre[(size_t)b * a.C + c];
      s = which == 5 ? fmaf(d, a.m[(size_t)b * a.C + c], s) : s + d;
    }
    (which == 5 ? a.g_gfc_w : a.g_gfc_b)[c] = s;
  }
}

// ---- GroupNorm backward.  CTA per (group, b).  dx = add + rstd*(dh - mean(dh) - xhat*mean(dh*xhat)) ----
__global__ void __launch_bounds__(TS_THREADS)
groupnorm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dg, int T, int C, int groups,
                     const float* __restrict__ gamma, const float* __restrict__ add, float* __restrict__ dy,
                     float* __restrict__ part /* [B][2][C] */) {
  __shared__ float s_red[32];
  const int cg = C / groups;
  const int grp = blockIdx.x, b = blockIdx.y;
  const int c0 = grp * cg;
  const int n = T * cg;
  const float* yb = y + (size_t)b * T * C;
  const float* db = dg + (size_t)b * T * C;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += TS_THREADS) s += yb[(size_t)(i / cg) * C + c0 + i % cg];
  const float mean = block_sum(s, s_red) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += TS_THREADS) {
    const float d = yb[(size_t)(i / cg) * C + c0 + i % cg] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.f / sqrtf(block_sum(q, s_red) / (float)n + TS_EPS);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n; i += TS_THREADS) {
    const size_t at = (size_t)(i / cg) * C + c0 + i % cg;
    const float dh = db[at] * gamma[c0 + i % cg];
    s1 += dh;
    s2 = fmaf(dh, (yb[at] - mean) * rstd, s2);
  }
  s1 = block_sum(s1, s_red) / (float)n;
  s2 = block_sum(s2, s_red) / (float)n;
  for (int i = threadIdx.x; i < n; i += TS_THREADS) {
    const size_t at = (size_t)(i / cg) * C + c0 + i % cg;
    const float xh = (yb[at] - mean) * rstd;
    const float dh = db[at] * gamma[c0 + i % cg];
    float v = rstd * (dh - s1 - xh * s2);
    if (add) v += add[(size_t)b * T * C + at];
    dy[(size_t)b * T * C + at] = v;
  }
  // per-(b, channel) parameter partials
  for (int cl = threadIdx.x; cl < cg; cl += TS_THREADS) {
    float a = 0.f, bb = 0.f;
    for (int t = 0; t < T; ++t) {
      const size_t at = (size_t)t * C + c0 + cl;
      a = fmaf(db[at], (yb[at] - mean) * rstd, a);
      bb += db[at];
    }
    part[((size_t)b * 2) * C + c0 + cl] = a;
    part[((size_t)b * 2 + 1) * C + c0 + cl] = bb;
  }
}

__global__ void gn_param_final_kernel(const float* __restrict__ part, int B, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int i = 0; i < B; ++i) {
    a += part[((size_t)i * 2) * C + c];
    b += part[((size_t)i * 2 + 1) * C + c];
  }
  dgamma[c] = a;
  dbeta[c] = b;
}

// ---- GELU ----
template <typename TO>
__global__ void gelu_fwd_kernel(const float* __restrict__ h, long long n, TO* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Elem<TO>::st(out + i, gelu_erf(h[i]));
}
template <typename TO>
__global__ void gelu_bwd_kernel(const float* __restrict__ h, const float* __restrict__ da, long long n, TO* __restrict__ dh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = h[i];
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  Elem<TO>::st(dh + i, da[i] * (cdf + x * pdf));
}

// ---- linear upsample (align_corners=True) backward: dx[b, i, c] = sum_t weight(t, i) * dxu[b, t, c] ----
__global__ void upsample_bwd_kernel(const float* __restrict__ dxu, int B, int tc, int T, int C, float* __restrict__ dx) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * tc * C) return;
  const int c = (int)(idx % C);
  const int i = (int)((idx / C) % tc);
  const int b = (int)(idx / ((long long)C * tc));
  const float scale = T > 1 ? (float)(tc - 1) / (float)(T - 1) : 0.f;
  float s = 0.f;
  for (int t = 0; t < T; ++t) {
    const float pos = scale * (float)t;
    int i0 = (int)pos;
    if (i0 > tc - 1) i0 = tc - 1;
    const int i1 = min(i0 + 1, tc - 1);
    const float l1 = pos - (float)i0, l0 = 1.f - l1;
    float wgt = 0.f;
    if (i0 == i) wgt += l0;
    if (i1 == i) wgt += l1;
    if (wgt != 0.f) s = fmaf(wgt, dxu[((size_t)b * T + t) * C + c], s);
  }
  dx[idx] = s;
}

template <typename TO>
__global__ void cast_kernel(const float* __restrict__ in, long long n, TO* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Elem<TO>::st(out + i, in[i]);
}

}  // namespace tdeed

using namespace tdeed;

extern "C" int tdeed_chan_ln_fwd(const float* x, int B, int t_in, int T, int C, const float* w, const float* b, float* xp,
                                 int* argmax, float* ln, float* stats, void* stream) {
  TDEED_REQUIRE(x && w && b && ln && stats && B > 0 && T > 0 && t_in >= T && C > 0, TDEED_ERR_SHAPE, "tdeed_chan_ln_fwd: bad arguments");
  chan_ln_fwd_kernel<<<ceil_div(B * T, TS_THREADS / 32), TS_THREADS, 0, (cudaStream_t)stream>>>(x, B, t_in, T, C, w, b, xp, argmax, ln, stats);
  return check_launch("tdeed_chan_ln_fwd");
}

extern "C" int tdeed_chan_ln_bwd(const float* xp, const float* stats, const float* dln, long long ld, int rows, int C,
                                 const float* w, const float* add, float* dx, float* dw, float* db, void* stream) {
  TDEED_REQUIRE(xp && stats && dln && w && dx && dw && db && rows > 0 && C > 0 && ld >= C, TDEED_ERR_SHAPE, "tdeed_chan_ln_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  chan_ln_bwd_kernel<<<ceil_div(rows, TS_THREADS / 32), TS_THREADS, 0, st>>>(xp, stats, dln, ld, rows, C, w, add, dx);
  int rc = check_launch("tdeed_chan_ln_bwd");
  if (rc) return rc;
  chan_ln_param_kernel<<<ceil_div(C, 128), 128, 0, st>>>(xp, stats, dln, ld, rows, C, dw, db);
  return check_launch("tdeed_chan_ln_bwd(param)");
}

extern "C" int tdeed_maxpool_bwd(const float* dxp, const int* argmax, int B, int t_in, int T, int C, const float* add, float* dx,
                                 void* stream) {
  TDEED_REQUIRE(dxp && argmax && dx && B > 0 && T > 0 && t_in >= T && C > 0, TDEED_ERR_SHAPE, "tdeed_maxpool_bwd: bad arguments");
  const long long n = (long long)B * t_in * C;
  maxpool_bwd_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(dxp, argmax, B, t_in, T, C, add, dx);
  return check_launch("tdeed_maxpool_bwd");
}

extern "C" long long tdeed_sgp_branch_bwd_workspace_floats(int B, int T, int C) { return 3LL * B * C + 2LL * B * T * C; }

extern "C" int tdeed_sgp_branch_bwd(const float* ln, long long ld_ln, const float* d_conv, const float* d_fc, const float* d_id,
                                    long long ld_g, int B, int T, int C, int ks, int up, const float* const* weights /*10*/,
                                    float* const* grads /*10*/, float* d_ln, float* workspace, void* stream) {
  TDEED_REQUIRE(ln && d_conv && d_fc && weights && grads && d_ln && workspace, TDEED_ERR_SHAPE, "tdeed_sgp_branch_bwd: null pointer");
  TDEED_REQUIRE(B > 0 && T > 0 && C > 0 && ks > 0 && up > 0 && (ks & 1) && (up & 1) && ld_ln >= C && ld_g >= C, TDEED_ERR_SHAPE,
                "tdeed_sgp_branch_bwd: bad shape B=%d T=%d C=%d ks=%d up=%d", B, T, C, ks, up);
  BranchArgs a;
  a.ln = ln; a.ld_ln = ld_ln; a.d_conv = d_conv; a.d_fc = d_fc; a.d_id = d_id; a.ld_g = ld_g;
  a.psi_w = weights[0]; a.psi_b = weights[1]; a.convw_w = weights[2]; a.convw_b = weights[3]; a.convkw_w = weights[4];
  a.convkw_b = weights[5]; a.fc_w = weights[6]; a.fc_b = weights[7]; a.gfc_w = weights[8]; a.gfc_b = weights[9];
  a.g_psi_w = grads[0]; a.g_psi_b = grads[1]; a.g_convw_w = grads[2]; a.g_convw_b = grads[3]; a.g_convkw_w = grads[4];
  a.g_convkw_b = grads[5]; a.g_fc_w = grads[6]; a.g_fc_b = grads[7]; a.g_gfc_w = grads[8]; a.g_gfc_b = grads[9];
  for (int i = 0; i < 10; ++i) TDEED_REQUIRE(weights[i] && grads[i], TDEED_ERR_SHAPE, "tdeed_sgp_branch_bwd: null weight/grad %d", i);
  a.B = B; a.T = T; a.C = C; a.ks = ks; a.up = up;
  a.m = workspace; a.dpre = a.m + (size_t)B * C; a.phi = a.dpre + (size_t)B * C;
  a.dcwk = a.phi + (size_t)B * C; a.dpsi = a.dcwk + (size_t)B * T * C;
  a.d_ln = d_ln;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * T * C;
  branch_k1<<<ceil_div(B * C, 128), 128, 0, st>>>(a);
  int rc = check_launch("tdeed_sgp_branch_bwd(k1)");
  if (rc) return rc;
  branch_k2<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(a);
  if ((rc = check_launch("tdeed_sgp_branch_bwd(k2)"))) return rc;
  branch_k3<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(a);
  if ((rc = check_launch("tdeed_sgp_branch_bwd(k3)"))) return rc;
  branch_k4<<<(unsigned)ceil_div_ll((long long)C * (2 * ks + up + 7), 128), 128, 0, st>>>(a);
  return check_launch("tdeed_sgp_branch_bwd(k4)");
}

extern "C" int tdeed_groupnorm_bwd(const float* y, const float* dg, int B, int T, int C, int groups, const float* gamma,
                                   const float* add, float* dy, float* dgamma, float* dbeta, float* workspace, void* stream) {
  TDEED_REQUIRE(y && dg && gamma && dy && dgamma && dbeta && workspace && B > 0 && T > 0 && C > 0 && groups > 0 && C % groups == 0,
                TDEED_ERR_SHAPE, "tdeed_groupnorm_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  groupnorm_bwd_kernel<<<dim3(groups, B), TS_THREADS, 0, st>>>(y, dg, T, C, groups, gamma, add, dy, workspace);
  int rc = check_launch("tdeed_groupnorm_bwd");
  if (rc) return rc;
  gn_param_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(workspace, B, C, dgamma, dbeta);
  return check_launch("tdeed_groupnorm_bwd(param)");
}

extern "C" int tdeed_gelu_fwd(const float* h, long long n, void* out, int out_dtype, void* stream) {
  TDEED_REQUIRE(h && out && n > 0, TDEED_ERR_SHAPE, "tdeed_gelu_fwd: bad arguments");
  const unsigned grid = (unsigned)ceil_div_ll(n, 256);
  if (out_dtype == TDEED_BF16) gelu_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(h, n, (__nv_bfloat16*)out);
  else if (out_dtype == TDEED_F32) gelu_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(h, n, (float*)out);
  else { set_error("tdeed_gelu_fwd: dtype %d", out_dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_gelu_fwd");
}

extern "C" int tdeed_gelu_bwd(const float* h, const float* da, long long n, void* dh, int out_dtype, void* stream) {
  TDEED_REQUIRE(h && da && dh && n > 0, TDEED_ERR_SHAPE, "tdeed_gelu_bwd: bad arguments");
  const unsigned grid = (unsigned)ceil_div_ll(n, 256);
  if (out_dtype == TDEED_BF16) gelu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(h, da, n, (__nv_bfloat16*)dh);
  else if (out_dtype == TDEED_F32) gelu_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(h, da, n, (float*)dh);
  else { set_error("tdeed_gelu_bwd: dtype %d", out_dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_gelu_bwd");
}

extern "C" int tdeed_upsample_bwd(const float* dxu, int B, int t_coarse, int T, int C, float* dx, void* stream) {
  TDEED_REQUIRE(dxu && dx && B > 0 && t_coarse > 0 && T >= t_coarse && C > 0, TDEED_ERR_SHAPE, "tdeed_upsample_bwd: bad arguments");
  const long long n = (long long)B * t_coarse * C;
  upsample_bwd_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, (cudaStream_t)stream>>>(dxu, B, t_coarse, T, C, dx);
  return check_launch("tdeed_upsample_bwd");
}

extern "C" int tdeed_cast_f32(const float* in, long long n, void* out, int out_dtype, void* stream) {
  TDEED_REQUIRE(in && out && n > 0, TDEED_ERR_SHAPE, "tdeed_cast_f32: bad arguments");
  const unsigned grid = (unsigned)ceil_div_ll(n, 256);
  if (out_dtype == TDEED_BF16) cast_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(in, n, (__nv_bfloat16*)out);
  else if (out_dtype == TDEED_F32) cast_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(in, n, (float*)out);
  else { set_error("tdeed_cast_f32: dtype %d", out_dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_cast_f32");
}
