// (7)(8)(9) SGP / SGP-Mixer token mixing and GroupNorm on [B, T, C] fp32 sequences.
// Reference: model/modules.py:159-186 (SGPBlock), :283-307 (SGPMixer), :348-363 (channel LayerNorm),
// :64,76 (AdaptiveMaxPool1d), :236,288 (linear upsample, align_corners=True).
//
// Round-2 design: the work is partitioned over B x T x C (round 1 used 16 x B CTAs running serial phases and reached
// 0.5-2 % of the HBM roofline).  Every pass streams whole rows / [rows x channels] tiles with coalesced accesses; the two
// clip-wide reductions (mean over T of the LayerNorm output for the phi gate, GroupNorm statistics over T x C/16) are
// two-level and deterministic: tiles write partial sums, consumers add them in a fixed order.
//
//   sgp_rowstats_kernel  warp per (pooled) row: LayerNorm mean / rstd over C, row kept in registers; the CTA's 32 rows also
//                        produce per-channel partial sums of the normalised values (times a row weight) -> phi partials
//   sgp_mix_kernel       tile [32 rows + halo] x [64 channels]: LN output staged in shared memory, both depthwise convolutions
//                        as register sliding windows (8 consecutive rows per thread), gating, residual -> y; per-channel
//                        partial (sum, sum of squares) of y in double -> GroupNorm partials
//   sgp_gnstats_kernel   (16 groups x B) tiny: GroupNorm mean / rstd from the partials
//   sgp_gnapply_kernel   g = GN(y) as the MLP's GEMM operand (bf16 | fp32), 16 bytes per thread
//   sgp_mixer_kernel     same tile scheme for SGPMixer: z = LN1(skip), u = upsample(LN2(x)); writes the 6C-wide concat
// convw and convkw read the same input, so their kernels are merged into one `up`-tap kernel.
// mean_T(LN(x))[c] = ln_w[c] * (sum_t nrm[t, c]) / T + ln_b[c]  with nrm = (x - mean_t) * rstd_t, and for the upsampled
// operand sum_t u[t, c] = sum_i coef_i * a[i, c] with coef_i = total interpolation weight of coarse row i: both gates come
// from the row pass without touching the sequence again.
#include "common.cuh"

namespace tdeed {

constexpr int SG_THREADS = 256;
constexpr int SG_GROUPS = 16;
constexpr float SG_EPS = 1e-5f;
constexpr int SG_RB = 32;          // rows per row-statistics CTA (4 per warp)
constexpr int SG_TT = 64;          // output rows per mixing tile
constexpr int SG_CB = 64;          // channels per mixing tile
constexpr int SG_RT = 8;           // consecutive rows per thread (sliding window)
constexpr int SG_MIX_THREADS = SG_CB * (SG_TT / SG_RT);   // 512: thread = (channel, group of SG_RT rows)
constexpr int SG_RG = SG_TT / SG_RT;                       // row groups per tile
constexpr int SG_MAXV = 8;         // float4 per lane per row: C <= 1024
constexpr int SG_MAXROWS = SG_TT + 2 * 64;                 // tile rows incl. halo: up <= 129

struct PoolWin { int s, e; };
__device__ inline PoolWin pool_window(int t, int t_in, int t_out) {   // AdaptiveMaxPool1d window
  PoolWin wdw;
  wdw.s = (t * t_in) / t_out;                        // t * t_in < 2^31 (host check)
  wdw.e = ((t + 1) * t_in + t_out - 1) / t_out;
  return wdw;
}
__device__ inline float pooled(const float* __restrict__ xb, int C, int c, PoolWin wdw) {
  float m = xb[(size_t)wdw.s * C + c];
  for (int r = wdw.s + 1; r < wdw.e; ++r) m = fmaxf(m, xb[(size_t)r * C + c]);
  return m;
}

// linear upsample (align_corners=True) source rows / weights of output row t
struct Lerp { int i0, i1; float l0, l1; };
__device__ inline Lerp lerp_src(int t, int tc, float scale) {
  Lerp s;
  const float real = scale * (float)t;
  s.i0 = (int)real;
  s.i1 = s.i0 + ((s.i0 < tc - 1) ? 1 : 0);
  s.l1 = fminf(fmaxf(real - (float)s.i0, 0.f), 1.f);
  s.l0 = 1.f - s.l1;
  return s;
}

// ---- pass 1: LayerNorm statistics per row + per-channel partial sums of the normalised rows ---------------------------
// grid (ceil(T / SG_RB), B).  stats: [B][T][2] (mean, rstd).  colpart: [B][nrb][C].  up_T > 0: rows are the COARSE rows of a
// mixer (weight = total linear-interpolation weight of the row when the sequence is upsampled to up_T rows); else weight 1.
template <int NV>                                         // float4 slots per lane: C <= 128 * NV
__global__ void __launch_bounds__(SG_THREADS)
sgp_rowstats_kernel(const float* __restrict__ x, int t_in, int T, int C, int up_T, float* __restrict__ stats,
                    float* __restrict__ colpart) {
  extern __shared__ __align__(16) float smem[];          // [8 warps][C] column partials of the CTA's rows
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int c4 = C >> 2;
  const float* xb = x + (size_t)b * t_in * C;
  // per-warp column accumulators live in shared memory (registers hold the row itself: C = 768 would need 48 more)
  for (int q = lane; q < c4; q += 32) *reinterpret_cast<float4*>(smem + (size_t)warp * C + 4 * q) = make_float4(0.f, 0.f, 0.f, 0.f);
  const float up_scale = (up_T > 1) ? (float)(T - 1) / (float)(up_T - 1) : 0.f;
  for (int rr = 0; rr < SG_RB / 8; ++rr) {
    const int t = blockIdx.x * SG_RB + rr * 8 + warp;           // warp-uniform
    if (t >= T) break;
    const PoolWin wdw = pool_window(t, t_in, T);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int q = lane + 32 * k;
      if (q < c4) {
        float4 m = *reinterpret_cast<const float4*>(xb + (size_t)wdw.s * C + 4 * q);
        for (int r = wdw.s + 1; r < wdw.e; ++r) {
          const float4 o = *reinterpret_cast<const float4*>(xb + (size_t)r * C + 4 * q);
          m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
        }
        v[k] = m;
        s += (m.x + m.y) + (m.z + m.w);
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float qs = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane + 32 * k < c4) {
        const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
        qs += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = 1.f / sqrtf(warp_sum(qs) / (float)C + SG_EPS);
    float wt = 1.f;
    if (up_T > 0) {                                        // total interpolation weight of coarse row t
      float acc = 0.f;
      for (int u = lane; u < up_T; u += 32) {
        const Lerp ls = lerp_src(u, T, up_scale);
        if (ls.i0 == t) acc += ls.l0;
        if (ls.i1 == t) acc += ls.l1;
      }
      wt = warp_sum(acc);
    }
    if (lane == 0) {
      stats[((size_t)b * T + t) * 2] = mean;
      stats[((size_t)b * T + t) * 2 + 1] = rstd;
    }
    const float f = rstd * wt;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane + 32 * k < c4) {
        float4* a = reinterpret_cast<float4*>(smem + (size_t)warp * C + 4 * (lane + 32 * k));
        float4 acc = *a;
        acc.x = fmaf(v[k].x - mean, f, acc.x);
        acc.y = fmaf(v[k].y - mean, f, acc.y);
        acc.z = fmaf(v[k].z - mean, f, acc.z);
        acc.w = fmaf(v[k].w - mean, f, acc.w);
        *a = acc;
      }
    }
  }
  __syncthreads();
  float* out = colpart + ((size_t)b * gridDim.x + blockIdx.x) * C;
  for (int c = threadIdx.x; c < C; c += SG_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) a += smem[(size_t)w8 * C + c];       // fixed order
    out[c] = a;
  }
}

// phi[c] = relu(gfc_w * (ln_w * S / T + ln_b) + gfc_b),  S = sum over the row-block partials (fixed order)
__device__ inline float phi_gate(const float* __restrict__ colpart_b, int nrb, int C, int c, int T, float ln_w, float ln_b,
                                 float gfc_w, float gfc_b) {
  float s = 0.f;
  for (int r = 0; r < nrb; ++r) s += colpart_b[(size_t)r * C + c];
  return fmaxf(fmaf(gfc_w, fmaf(ln_w, s / (float)T, ln_b), gfc_b), 0.f);
}

// depthwise temporal convolution of SG_RT consecutive rows of one channel: acc[r] += sum_k w[k] * tile[row0 + r + k] (the
// tile is zero outside the sequence, so zero padding is implicit).  w: [k][SG_CB] in shared memory, tile: [rows][SG_CB].
// Taps are consumed in blocks of SG_RT with the 2*SG_RT-1 rows they touch held in registers (static indices: 16 shared-memory
// reads per 64 FMAs); the remaining ntap % SG_RT taps use a one-row-per-tap sliding window.
__device__ __forceinline__ void dw_window(const float* __restrict__ tile_col, const float* __restrict__ w_col, int ntap,
                                          float (&acc)[SG_RT]) {
  int k = 0;
  if (ntap >= SG_RT) {
    float v[2 * SG_RT - 1];
#pragma unroll
    for (int i = 0; i < SG_RT - 1; ++i) v[SG_RT + i] = tile_col[i * SG_CB];          // rows 0 .. RT-2 wait in the upper half
    for (; k + SG_RT <= ntap; k += SG_RT) {
#pragma unroll
      for (int i = 0; i < SG_RT - 1; ++i) v[i] = v[SG_RT + i];                         // rows k .. k+RT-2
#pragma unroll
      for (int i = SG_RT - 1; i < 2 * SG_RT - 1; ++i) v[i] = tile_col[(k + i) * SG_CB];   // rows k+RT-1 .. k+2RT-2
#pragma unroll
      for (int j = 0; j < SG_RT; ++j) {
        const float wk = w_col[(k + j) * SG_CB];
#pragma unroll
        for (int r = 0; r < SG_RT; ++r) acc[r] = fmaf(wk, v[r + j], acc[r]);
      }
    }
  }
  if (k < ntap) {
    float win[SG_RT];
#pragma unroll
    for (int r = 0; r < SG_RT - 1; ++r) win[r + 1] = tile_col[(k + r) * SG_CB];
    const float* next = tile_col + (size_t)(k + SG_RT - 1) * SG_CB;
    for (; k < ntap; ++k, next += SG_CB) {
#pragma unroll
      for (int r = 0; r < SG_RT - 1; ++r) win[r] = win[r + 1];
      win[SG_RT - 1] = *next;
      const float wk = w_col[k * SG_CB];
#pragma unroll
      for (int r = 0; r < SG_RT; ++r) acc[r] = fmaf(wk, win[r], acc[r]);
    }
  }
}

struct SgpW {
  tdeed_sgp_weights w;
};

// ---- pass 2 (SGP block): grid (ceil(T / SG_TT), ceil(C / SG_CB), B), SG_MIX_THREADS threads ------------------------------
__global__ void __launch_bounds__(SG_MIX_THREADS)
sgp_mix_kernel(const float* __restrict__ x, int t_in, int T, int C, int ks, int up, SgpW W, const float* __restrict__ stats,
               const float* __restrict__ colpart, int nrb, float* __restrict__ y, double* __restrict__ gnpart) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_ws[SG_MAXROWS], s_we[SG_MAXROWS];         // pooling window of every staged row (empty = outside [0, T))
  __shared__ float s_rm[SG_MAXROWS], s_rr[SG_MAXROWS];       // LayerNorm mean / rstd of every staged row
  const tdeed_sgp_weights& w = W.w;
  const int h = up / 2, hp = ks / 2;
  const int rows = SG_TT + 2 * h;
  float* s_ln = smem;                               // [rows][CB]   LN output, zero outside [0, T)
  float* s_xp = s_ln + (size_t)rows * SG_CB;        // [TT][CB]     pooled input of the tile's own rows
  float* s_wm = s_xp + SG_TT * SG_CB;               // [up][CB]     convkw with convw folded into the centre taps
  float* s_wp = s_wm + (size_t)up * SG_CB;          // [ks][CB]     psi
  float* s_phi = s_wp + (size_t)ks * SG_CB;         // [CB]
  double* s_gn = reinterpret_cast<double*>(s_phi + SG_CB);   // [RG][CB][2]
  const int b = blockIdx.z, c0 = blockIdx.y * SG_CB, t0 = blockIdx.x * SG_TT;
  const int cl = threadIdx.x % SG_CB, tr = threadIdx.x / SG_CB;          // SG_RG row groups of SG_RT rows
  const int c = c0 + cl;
  const bool cok = c < C;
  const float* xb = x + (size_t)b * t_in * C;
  const float* st = stats + (size_t)b * T * 2;

  for (int r = threadIdx.x; r < rows; r += SG_MIX_THREADS) {              // per-row table: one thread per row
    const int t = t0 - h + r;
    int ws = 0, we = 0;
    float m = 0.f, rs = 0.f;
    if (t >= 0 && t < T) {
      if (t_in == T) { ws = t; we = t + 1; }
      else { const PoolWin pw = pool_window(t, t_in, T); ws = pw.s; we = pw.e; }
      m = st[2 * t];
      rs = st[2 * t + 1];
    }
    s_ws[r] = ws; s_we[r] = we; s_rm[r] = m; s_rr[r] = rs;
  }
  const float lnw = cok ? w.ln_w[c] : 0.f, lnb = cok ? w.ln_b[c] : 0.f;
  for (int k = tr; k < up; k += SG_RG) {
    float v = 0.f;
    if (cok) {
      v = w.convkw_w[(size_t)c * up + k];
      const int kk = k - (h - hp);
      if (kk >= 0 && kk < ks) v += w.convw_w[(size_t)c * ks + kk];
    }
    s_wm[k * SG_CB + cl] = v;
  }
  for (int k = tr; k < ks; k += SG_RG) s_wp[k * SG_CB + cl] = cok ? w.psi_w[(size_t)c * ks + k] : 0.f;
  if (tr == 0) s_phi[cl] = cok ? phi_gate(colpart + (size_t)b * nrb * C, nrb, C, c, T, lnw, lnb, w.gfc_w[c], w.gfc_b[c]) : 0.f;
  __syncthreads();
  // (32-bit offsets: a clip has < 2^31 elements.  The first version of this loop — 64-bit row * C products and the generic
  // pooling loop unrolled by the compiler — was 385 SASS instructions per row and 3/4 of the kernel's issue slots: ncu r2.)
  {
    const float* xc = xb + c;
    float* dst = s_ln + tr * SG_CB + cl;
#pragma unroll 2
    for (int r = tr; r < rows; r += SG_RG, dst += SG_RG * SG_CB) {
      const int ws = s_ws[r], nw = s_we[r] - ws;
      float ln = 0.f;
      if (cok && nw > 0) {
        const float* px = xc + ws * C;
        float xv = px[0];
        if (nw > 1) {
          xv = fmaxf(xv, px[C]);
#pragma unroll 1
          for (int q = 2; q < nw; ++q) xv = fmaxf(xv, px[q * C]);
        }
        ln = fmaf((xv - s_rm[r]) * s_rr[r], lnw, lnb);
        if (r >= h && r < h + SG_TT) s_xp[(r - h) * SG_CB + cl] = xv;
      }
      *dst = ln;
    }
  }
  __syncthreads();

  const int rl = tr * SG_RT;                        // first local row of this thread
  float am[SG_RT], ap[SG_RT];
  const float bm = cok ? w.convw_b[c] + w.convkw_b[c] : 0.f, bp = cok ? w.psi_b[c] : 0.f;
#pragma unroll
  for (int r = 0; r < SG_RT; ++r) { am[r] = bm; ap[r] = bp; }
  dw_window(s_ln + (size_t)rl * SG_CB + cl, s_wm + cl, up, am);                       // taps t-h .. t+h
  dw_window(s_ln + (size_t)(rl + h - hp) * SG_CB + cl, s_wp + cl, ks, ap);            // taps t-hp .. t+hp
  float* yb = y + (size_t)b * T * C;
  // GroupNorm partials: fp32 over the thread's 8 rows, double from there on (fp64 issue is scarce on this part: per-element
  // double arithmetic made the whole kernel FP64-bound — 250 issue slots per element, ncu r2)
  float fsum = 0.f, fsq = 0.f;
  if (cok) {
    const float fcw = w.fc_w[c], fcb = w.fc_b[c], phi = s_phi[cl];
#pragma unroll
    for (int r = 0; r < SG_RT; ++r) {
      const int t = t0 + rl + r;
      if (t < T) {
        const float ln = s_ln[(size_t)(h + rl + r) * SG_CB + cl];
        const float yv = s_xp[(rl + r) * SG_CB + cl] + (fmaf(fcw, ln, fcb) * phi + am[r] * ap[r] + ln);
        yb[t * C + c] = yv;
        fsum += yv;
        fsq = fmaf(yv, yv, fsq);
      }
    }
  }
  s_gn[(tr * SG_CB + cl) * 2] = (double)fsum;
  s_gn[(tr * SG_CB + cl) * 2 + 1] = (double)fsq;
  __syncthreads();
  if (tr == 0 && cok) {
    double a = 0.0, q = 0.0;
#pragma unroll
    for (int g = 0; g < SG_RG; ++g) { a += s_gn[(g * SG_CB + cl) * 2]; q += s_gn[(g * SG_CB + cl) * 2 + 1]; }
    double* o = gnpart + (((size_t)b * gridDim.x + blockIdx.x) * C + c) * 2;
    o[0] = a;
    o[1] = q;
  }
}

// ---- pass 3: GroupNorm statistics from the partials.  grid (groups, B), one warp-sized reduction tree per CTA ----------
__global__ void __launch_bounds__(SG_THREADS)
sgp_gnstats_kernel(const double* __restrict__ gnpart, int ntt, int T, int C, int groups, float* __restrict__ gstats) {
  __shared__ double s_a[SG_THREADS], s_q[SG_THREADS];
  const int cg = C / groups, g = blockIdx.x, b = blockIdx.y;
  double a = 0.0, q = 0.0;
  const int n = ntt * cg;
  for (int i = threadIdx.x; i < n; i += SG_THREADS) {
    const int tile = i / cg, cl = i - tile * cg;
    const double* p = gnpart + (((size_t)b * ntt + tile) * C + g * cg + cl) * 2;
    a += p[0];
    q += p[1];
  }
  s_a[threadIdx.x] = a;
  s_q[threadIdx.x] = q;
  __syncthreads();
  for (int o = SG_THREADS / 2; o > 0; o >>= 1) {            // fixed tree
    if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_q[threadIdx.x] += s_q[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double cnt = (double)T * (double)cg;
    const double mean = s_a[0] / cnt;
    double var = s_q[0] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    gstats[((size_t)b * groups + g) * 2] = (float)mean;
    gstats[((size_t)b * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)SG_EPS));
  }
}

__device__ inline void st_cat(void* cat, int dtype, size_t o, float v) {
  if (dtype == TDEED_F32) reinterpret_cast<float*>(cat)[o] = v;
  else reinterpret_cast<__nv_bfloat16*>(cat)[o] = __float2bfloat16_rn(v);
}

// ---- pass 4: g = (y - mean_g) * rstd_g * gamma + beta, 4 channels per thread.  grid (ceil(T * C / 4 / 256), B) ----------
__global__ void __launch_bounds__(SG_THREADS)
sgp_gnapply_kernel(const float* __restrict__ y, int T, int C, int groups, const float* __restrict__ gstats,
                   const float* __restrict__ gamma, const float* __restrict__ beta, void* __restrict__ out, int out_dtype) {
  __shared__ float s_g[2 * 64];
  const int b = blockIdx.y;
  if (threadIdx.x < 2 * groups) s_g[threadIdx.x] = gstats[(size_t)b * groups * 2 + threadIdx.x];
  __syncthreads();
  const int c4n = C >> 2, cg = C / groups;
  const int i = blockIdx.x * SG_THREADS + threadIdx.x;
  if (i >= T * c4n) return;
  const int t = i / c4n;
  const int c = (i - t * c4n) * 4;
  const size_t base = ((size_t)b * T + t) * C + c;
  const float4 v = *reinterpret_cast<const float4*>(y + base);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
  const float4 be = *reinterpret_cast<const float4*>(beta + c);
  const float in[4] = {v.x, v.y, v.z, v.w}, gm[4] = {ga.x, ga.y, ga.z, ga.w}, bt[4] = {be.x, be.y, be.z, be.w};
  const int g0 = c / cg;
  const int edge = (g0 + 1) * cg;                   // first channel of the next group (4 channels span at most 2 groups: cg >= 4)
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int g = g0 + ((c + j) >= edge ? 1 : 0);
    o[j] = fmaf((in[j] - s_g[2 * g]) * s_g[2 * g + 1], gm[j], bt[j]);
  }
  if (out_dtype == TDEED_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + base) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
    uint2 pk;
    pk.x = pack_bf16x2(o[0], o[1]);
    pk.y = pack_bf16x2(o[2], o[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + base) = pk;
  }
}

struct MixW {
  tdeed_mixer_weights w;
};

// ---- SGPMixer token mixing: grid (ceil(T / SG_TT), ceil(C / SG_CB), B) -------------------------------------------------
// z = LN1(skip) [T rows], u = linear upsample of LN2(x) [tc rows] to T rows; cat = [o1, o2, o3, o4, z, u] (6C wide).
__global__ void __launch_bounds__(SG_MIX_THREADS)
sgp_mixer_kernel(const float* __restrict__ xc, const float* __restrict__ skip, int tc, int T, int C, int ks, int up, MixW W,
                 const float* __restrict__ stats_z, const float* __restrict__ colpart_z, int nrb_z,
                 const float* __restrict__ stats_x, const float* __restrict__ colpart_x, int nrb_x,
                 void* __restrict__ cat, int cat_dtype) {
  extern __shared__ __align__(16) float smem[];
  const tdeed_mixer_weights& w = W.w;
  const int h = up / 2, hp = ks / 2;
  const int rows = SG_TT + 2 * h;
  float* s_z = smem;                                // [rows][CB]
  float* s_u = s_z + (size_t)rows * SG_CB;          // [rows][CB]
  float* s_m1 = s_u + (size_t)rows * SG_CB;         // [up][CB]
  float* s_m2 = s_m1 + (size_t)up * SG_CB;
  float* s_p1 = s_m2 + (size_t)up * SG_CB;          // [ks][CB]
  float* s_p2 = s_p1 + (size_t)ks * SG_CB;
  float* s_phi = s_p2 + (size_t)ks * SG_CB;         // [2][CB]
  const int b = blockIdx.z, c0 = blockIdx.y * SG_CB, t0 = blockIdx.x * SG_TT;
  const int cl = threadIdx.x % SG_CB, tr = threadIdx.x / SG_CB;
  const int c = c0 + cl;
  const bool cok = c < C;
  const float* zb = skip + (size_t)b * T * C;
  const float* xb = xc + (size_t)b * tc * C;
  const float* sz = stats_z + (size_t)b * T * 2;
  const float* sx = stats_x + (size_t)b * tc * 2;
  const float l1w = cok ? w.ln1_w[c] : 0.f, l1b = cok ? w.ln1_b[c] : 0.f;
  const float l2w = cok ? w.ln2_w[c] : 0.f, l2b = cok ? w.ln2_b[c] : 0.f;

  __shared__ int s_i0[SG_MAXROWS], s_i1[SG_MAXROWS];         // upsample sources of every staged row (i0 < 0: outside [0, T))
  __shared__ float s_l1[SG_MAXROWS];
  __shared__ float s_zm[SG_MAXROWS], s_zr[SG_MAXROWS], s_x0m[SG_MAXROWS], s_x0r[SG_MAXROWS], s_x1m[SG_MAXROWS], s_x1r[SG_MAXROWS];
  const float scale = (T > 1) ? (float)(tc - 1) / (float)(T - 1) : 0.f;
  for (int r = threadIdx.x; r < rows; r += SG_MIX_THREADS) {
    const int t = t0 - h + r;
    int i0 = -1, i1 = 0;
    float l1 = 0.f, zm = 0.f, zr = 0.f, x0m = 0.f, x0r = 0.f, x1m = 0.f, x1r = 0.f;
    if (t >= 0 && t < T) {
      const Lerp ls = lerp_src(t, tc, scale);
      i0 = ls.i0; i1 = ls.i1; l1 = ls.l1;
      zm = sz[2 * t]; zr = sz[2 * t + 1];
      x0m = sx[2 * i0]; x0r = sx[2 * i0 + 1];
      x1m = sx[2 * i1]; x1r = sx[2 * i1 + 1];
    }
    s_i0[r] = i0; s_i1[r] = i1; s_l1[r] = l1;
    s_zm[r] = zm; s_zr[r] = zr; s_x0m[r] = x0m; s_x0r[r] = x0r; s_x1m[r] = x1m; s_x1r[r] = x1r;
  }
  for (int k = tr; k < up; k += SG_RG) {
    float v1 = 0.f, v2 = 0.f;
    if (cok) {
      v1 = w.convkw1_w[(size_t)c * up + k];
      v2 = w.convkw2_w[(size_t)c * up + k];
      const int kk = k - (h - hp);
      if (kk >= 0 && kk < ks) { v1 += w.convw1_w[(size_t)c * ks + kk]; v2 += w.convw2_w[(size_t)c * ks + kk]; }
    }
    s_m1[k * SG_CB + cl] = v1;
    s_m2[k * SG_CB + cl] = v2;
  }
  for (int k = tr; k < ks; k += SG_RG) {
    s_p1[k * SG_CB + cl] = cok ? w.psi1_w[(size_t)c * ks + k] : 0.f;
    s_p2[k * SG_CB + cl] = cok ? w.psi2_w[(size_t)c * ks + k] : 0.f;
  }
  if (tr == 0) s_phi[cl] = cok ? phi_gate(colpart_z + (size_t)b * nrb_z * C, nrb_z, C, c, T, l1w, l1b, w.gfc1_w[c], w.gfc1_b[c]) : 0.f;
  if (tr == 1) s_phi[SG_CB + cl] = cok ? phi_gate(colpart_x + (size_t)b * nrb_x * C, nrb_x, C, c, T, l2w, l2b, w.gfc2_w[c], w.gfc2_b[c]) : 0.f;
  __syncthreads();
  {
    const float* zc = zb + c;
    const float* xcc = xb + c;
#pragma unroll 2
    for (int r = tr; r < rows; r += SG_RG) {
      const int i0 = s_i0[r];
      float zv = 0.f, uv = 0.f;
      if (cok && i0 >= 0) {
        const int t = t0 - h + r;
        zv = fmaf((zc[t * C] - s_zm[r]) * s_zr[r], l1w, l1b);
        const float a0 = fmaf((xcc[i0 * C] - s_x0m[r]) * s_x0r[r], l2w, l2b);
        const float a1 = fmaf((xcc[s_i1[r] * C] - s_x1m[r]) * s_x1r[r], l2w, l2b);
        const float l1 = s_l1[r];
        uv = (1.f - l1) * a0 + l1 * a1;
      }
      s_z[r * SG_CB + cl] = zv;
      s_u[r * SG_CB + cl] = uv;
    }
  }
  __syncthreads();

  const int rl = tr * SG_RT;
  float am[SG_RT], ap[SG_RT];
  const size_t ldc = (size_t)6 * C;
  // operand 1: skip branch
  {
    const float bm = cok ? w.convw1_b[c] + w.convkw1_b[c] : 0.f, bp = cok ? w.psi1_b[c] : 0.f;
#pragma unroll
    for (int r = 0; r < SG_RT; ++r) { am[r] = bm; ap[r] = bp; }
    dw_window(s_z + (size_t)rl * SG_CB + cl, s_m1 + cl, up, am);
    dw_window(s_z + (size_t)(rl + h - hp) * SG_CB + cl, s_p1 + cl, ks, ap);
    if (cok) {
      const float fcw = w.fc1_w[c], fcb = w.fc1_b[c], phi = s_phi[cl];
#pragma unroll
      for (int r = 0; r < SG_RT; ++r) {
        const int t = t0 + rl + r;
        if (t < T) {
          const float zv = s_z[(size_t)(h + rl + r) * SG_CB + cl];
          const size_t row = ((size_t)b * T + t) * ldc + c;
          st_cat(cat, cat_dtype, row, am[r] * ap[r]);
          st_cat(cat, cat_dtype, row + 2 * (size_t)C, fmaf(fcw, zv, fcb) * phi);
          st_cat(cat, cat_dtype, row + 4 * (size_t)C, zv);
        }
      }
    }
  }
  // operand 2: upsampled coarse branch
  {
    const float bm = cok ? w.convw2_b[c] + w.convkw2_b[c] : 0.f, bp = cok ? w.psi2_b[c] : 0.f;
#pragma unroll
    for (int r = 0; r < SG_RT; ++r) { am[r] = bm; ap[r] = bp; }
    dw_window(s_u + (size_t)rl * SG_CB + cl, s_m2 + cl, up, am);
    dw_window(s_u + (size_t)(rl + h - hp) * SG_CB + cl, s_p2 + cl, ks, ap);
    if (cok) {
      const float fcw = w.fc2_w[c], fcb = w.fc2_b[c], phi = s_phi[SG_CB + cl];
#pragma unroll
      for (int r = 0; r < SG_RT; ++r) {
        const int t = t0 + rl + r;
        if (t < T) {
          const float uv = s_u[(size_t)(h + rl + r) * SG_CB + cl];
          const size_t row = ((size_t)b * T + t) * ldc + c;
          st_cat(cat, cat_dtype, row + C, am[r] * ap[r]);
          st_cat(cat, cat_dtype, row + 3 * (size_t)C, fmaf(fcw, uv, fcb) * phi);
          st_cat(cat, cat_dtype, row + 5 * (size_t)C, uv);
        }
      }
    }
  }
}

// ---- stand-alone GroupNorm (after the mixer's concat_fc GEMM): partial sums per 32-row tile, then the shared passes 3 / 4 ----
__global__ void __launch_bounds__(SG_MIX_THREADS)
sgp_gnpart_kernel(const float* __restrict__ x, int T, int C, double* __restrict__ gnpart) {
  __shared__ double s_gn[SG_RG * SG_CB * 2];
  const int b = blockIdx.z, c0 = blockIdx.y * SG_CB, t0 = blockIdx.x * SG_TT;
  const int cl = threadIdx.x % SG_CB, tr = threadIdx.x / SG_CB;
  const int c = c0 + cl;
  float a = 0.f, q = 0.f;
  if (c < C) {
#pragma unroll
    for (int r = 0; r < SG_RT; ++r) {
      const int t = t0 + tr * SG_RT + r;
      if (t < T) {
        const float v = x[((size_t)b * T + t) * C + c];
        a += v;
        q = fmaf(v, v, q);
      }
    }
  }
  s_gn[(tr * SG_CB + cl) * 2] = (double)a;
  s_gn[(tr * SG_CB + cl) * 2 + 1] = (double)q;
  __syncthreads();
  if (tr == 0 && c < C) {
    double sa = 0.0, sq = 0.0;
#pragma unroll
    for (int g = 0; g < SG_RG; ++g) { sa += s_gn[(g * SG_CB + cl) * 2]; sq += s_gn[(g * SG_CB + cl) * 2 + 1]; }
    double* o = gnpart + (((size_t)b * gridDim.x + blockIdx.x) * C + c) * 2;
    o[0] = sa;
    o[1] = sq;
  }
}

static int set_smem(const void* fn, size_t smem, const char* what, size_t* cur) {
  if (smem > *cur) {
    TDEED_REQUIRE(smem <= 227 * 1024, TDEED_ERR_UNSUPPORTED, "%s: needs %zu B of shared memory", what, smem);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    *cur = smem;
  }
  return TDEED_OK;
}

static inline long long align4(long long v) { return (v + 3) / 4 * 4; }

// workspace carving (floats; every region 16-byte aligned).  gn: [B][ntt][C][2] doubles = 4 floats per entry.
struct SgpWs {
  long long stats, colpart, gnpart, gstats, total;
  int nrb, ntt;
};
static SgpWs sgp_ws(int B, int T, int C) {
  SgpWs s;
  s.nrb = ceil_div(T, SG_RB);
  s.ntt = ceil_div(T, SG_TT);
  s.stats = 0;
  s.colpart = align4(2LL * B * T);
  s.gnpart = s.colpart + align4((long long)B * s.nrb * C);
  s.gstats = s.gnpart + (long long)B * s.ntt * C * 4;
  s.total = s.gstats + align4(2LL * B * SG_GROUPS);
  return s;
}

static int launch_rowstats(const float* x, int B, int t_in, int T, int C, int up_T, float* stats, float* colpart, cudaStream_t st,
                           const char* what) {
  const size_t smem = (size_t)8 * C * sizeof(float);           // <= 32 KB
  const dim3 grid(ceil_div(T, SG_RB), B);
  const int nv = ceil_div(C, 128);
  if (nv <= 2) sgp_rowstats_kernel<2><<<grid, SG_THREADS, smem, st>>>(x, t_in, T, C, up_T, stats, colpart);
  else if (nv <= 3) sgp_rowstats_kernel<3><<<grid, SG_THREADS, smem, st>>>(x, t_in, T, C, up_T, stats, colpart);
  else if (nv <= 4) sgp_rowstats_kernel<4><<<grid, SG_THREADS, smem, st>>>(x, t_in, T, C, up_T, stats, colpart);
  else if (nv <= 6) sgp_rowstats_kernel<6><<<grid, SG_THREADS, smem, st>>>(x, t_in, T, C, up_T, stats, colpart);
  else sgp_rowstats_kernel<SG_MAXV><<<grid, SG_THREADS, smem, st>>>(x, t_in, T, C, up_T, stats, colpart);
  return check_launch(what);
}

static int gn_finish(const float* y, int B, int T, int C, int groups, double* gnpart, int ntt, float* gstats, const float* gamma,
                     const float* beta, void* out, int out_dtype, cudaStream_t st, const char* what) {
  sgp_gnstats_kernel<<<dim3(groups, B), SG_THREADS, 0, st>>>(gnpart, ntt, T, C, groups, gstats);
  int rc = check_launch(what);
  if (rc) return rc;
  sgp_gnapply_kernel<<<dim3((unsigned)ceil_div(T * (C / 4), SG_THREADS), B), SG_THREADS, 0, st>>>(y, T, C, groups, gstats, gamma, beta,
                                                                                                 out, out_dtype);
  return check_launch(what);
}

}  // namespace tdeed

extern "C" long long tdeed_sgp_mix_workspace_floats(int B, int t_out, int C) { return tdeed::sgp_ws(B, t_out, C).total; }

extern "C" int tdeed_sgp_mix_fwd(const float* x, int B, int t_in, int t_out, int C, int ks, int up,
                                 const tdeed_sgp_weights* w_host, float* workspace, float* y, void* g, int g_dtype, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && w_host && workspace && y && g, TDEED_ERR_SHAPE, "tdeed_sgp_mix_fwd: null pointer");
  TDEED_REQUIRE(B > 0 && B <= 65535 && t_out > 0 && t_in >= t_out && C % SG_GROUPS == 0 && C <= 128 * SG_MAXV && ks % 2 == 1 &&
                up % 2 == 1 && up >= ks && up <= 129 && (long long)t_in * t_out < (1LL << 30), TDEED_ERR_SHAPE,
                "tdeed_sgp_mix_fwd: bad shape B=%d t_in=%d t_out=%d C=%d ks=%d up=%d", B, t_in, t_out, C, ks, up);
  TDEED_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, TDEED_ERR_SHAPE,
                "tdeed_sgp_mix_fwd: x and workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const SgpWs ws = sgp_ws(B, t_out, C);
  float* stats = workspace + ws.stats;
  float* colpart = workspace + ws.colpart;
  double* gnpart = reinterpret_cast<double*>(workspace + ws.gnpart);
  float* gstats = workspace + ws.gstats;
  static size_t cur_mix = 0;      // 0: always opt in (static + dynamic shared memory can exceed 48 KB even when dynamic alone does not)
  int rc = launch_rowstats(x, B, t_in, t_out, C, 0, stats, colpart, st, "tdeed_sgp_mix_fwd(rowstats)");
  if (rc) return rc;
  const size_t smem_mix = ((size_t)(SG_TT + 2 * (up / 2)) * SG_CB + (size_t)SG_TT * SG_CB + (size_t)(up + ks + 1) * SG_CB) * sizeof(float) +
                          (size_t)SG_RG * SG_CB * 2 * sizeof(double);
  rc = set_smem((const void*)sgp_mix_kernel, smem_mix, "tdeed_sgp_mix_fwd", &cur_mix);
  if (rc) return rc;
  SgpW W{*w_host};
  sgp_mix_kernel<<<dim3(ws.ntt, ceil_div(C, SG_CB), B), SG_MIX_THREADS, smem_mix, st>>>(x, t_in, t_out, C, ks, up, W, stats, colpart,
                                                                                        ws.nrb, y, gnpart);
  rc = check_launch("tdeed_sgp_mix_fwd(mix)");
  if (rc) return rc;
  return gn_finish(y, B, t_out, C, SG_GROUPS, gnpart, ws.ntt, gstats, w_host->gn_w, w_host->gn_b, g, g_dtype, st, "tdeed_sgp_mix_fwd(gn)");
}

extern "C" long long tdeed_sgp_mixer_workspace_floats(int B, int t_coarse, int T, int C) {
  using namespace tdeed;
  return align4(2LL * B * T) + align4((long long)B * ceil_div(T, SG_RB) * C) + align4(2LL * B * t_coarse) +
         align4((long long)B * ceil_div(t_coarse, SG_RB) * C);
}

extern "C" int tdeed_sgp_mixer_mix_fwd(const float* x_coarse, const float* skip, int B, int t_coarse, int T, int C,
                                       int ks, int up, const tdeed_mixer_weights* w_host, float* workspace, void* cat,
                                       int cat_dtype, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x_coarse && skip && w_host && workspace && cat, TDEED_ERR_SHAPE, "tdeed_sgp_mixer_mix_fwd: null pointer");
  TDEED_REQUIRE(B > 0 && B <= 65535 && T > 0 && t_coarse > 0 && C % SG_GROUPS == 0 && C <= 128 * SG_MAXV && ks % 2 == 1 && up % 2 == 1 &&
                up >= ks && up <= 129, TDEED_ERR_SHAPE, "tdeed_sgp_mixer_mix_fwd: bad shape B=%d tc=%d T=%d C=%d", B, t_coarse, T, C);
  TDEED_REQUIRE(((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(x_coarse) | reinterpret_cast<uintptr_t>(skip)) & 15) == 0,
                TDEED_ERR_SHAPE, "tdeed_sgp_mixer_mix_fwd: inputs and workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int nrb_z = ceil_div(T, SG_RB), nrb_x = ceil_div(t_coarse, SG_RB);
  float* stats_z = workspace;
  float* colpart_z = stats_z + align4(2LL * B * T);
  float* stats_x = colpart_z + align4((long long)B * nrb_z * C);
  float* colpart_x = stats_x + align4(2LL * B * t_coarse);
  static size_t cur_mix = 0;      // 0: always opt in (static + dynamic shared memory can exceed 48 KB even when dynamic alone does not)
  int rc = launch_rowstats(skip, B, T, T, C, 0, stats_z, colpart_z, st, "tdeed_sgp_mixer_mix_fwd(rowstats z)");
  if (rc) return rc;
  rc = launch_rowstats(x_coarse, B, t_coarse, t_coarse, C, T, stats_x, colpart_x, st, "tdeed_sgp_mixer_mix_fwd(rowstats x)");
  if (rc) return rc;
  const size_t smem_mix = ((size_t)2 * (SG_TT + 2 * (up / 2)) * SG_CB + (size_t)(2 * up + 2 * ks + 2) * SG_CB) * sizeof(float);
  rc = set_smem((const void*)sgp_mixer_kernel, smem_mix, "tdeed_sgp_mixer_mix_fwd", &cur_mix);
  if (rc) return rc;
  MixW W{*w_host};
  sgp_mixer_kernel<<<dim3(ceil_div(T, SG_TT), ceil_div(C, SG_CB), B), SG_MIX_THREADS, smem_mix, st>>>(
      x_coarse, skip, t_coarse, T, C, ks, up, W, stats_z, colpart_z, nrb_z, stats_x, colpart_x, nrb_x, cat, cat_dtype);
  return check_launch("tdeed_sgp_mixer_mix_fwd");
}

extern "C" long long tdeed_groupnorm_workspace_floats(int B, int T, int C, int groups) {
  using namespace tdeed;
  return (long long)B * ceil_div(T, SG_TT) * C * 4 + align4(2LL * B * groups);
}

extern "C" int tdeed_groupnorm_fwd(const float* x, int B, int T, int C, int groups, const float* gamma, const float* beta,
                                   float* workspace, void* out, int out_dtype, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && gamma && beta && workspace && out, TDEED_ERR_SHAPE, "tdeed_groupnorm_fwd: null pointer");
  TDEED_REQUIRE(B > 0 && B <= 65535 && T > 0 && groups > 0 && C % groups == 0 && C % 4 == 0, TDEED_ERR_SHAPE,
                "tdeed_groupnorm_fwd: bad shape B=%d T=%d C=%d groups=%d", B, T, C, groups);
  TDEED_REQUIRE(((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, TDEED_ERR_SHAPE,
                "tdeed_groupnorm_fwd: x and workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int ntt = ceil_div(T, SG_TT);
  double* gnpart = reinterpret_cast<double*>(workspace);
  float* gstats = workspace + (long long)B * ntt * C * 4;
  sgp_gnpart_kernel<<<dim3(ntt, ceil_div(C, SG_CB), B), SG_MIX_THREADS, 0, st>>>(x, T, C, gnpart);
  int rc = check_launch("tdeed_groupnorm_fwd(partials)");
  if (rc) return rc;
  return gn_finish(x, B, T, C, groups, gnpart, ntt, gstats, gamma, beta, out, out_dtype, st, "tdeed_groupnorm_fwd");
}
