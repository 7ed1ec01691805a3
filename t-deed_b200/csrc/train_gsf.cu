// Backward of the Gate-Shift(-Fuse) module with training-mode BatchNorm3d (model/impl/gsf.py:38-93, gsm.py:89-116,
// model/shift.py:89-93).  Forward (per clip, channels ch of group g, frame t, pixel p), as computed by gsf.cu:
//   z = relu(bn(x));  pre = conv3d(z) + b;  gate = tanh(pre);  y = gate*x;  r = x - y;  ys = shift_t(y)
//   GSF: P0 = mean_p ys, P1 = mean_p r;  A = conv2d_{2->1,3x3}([P0,P1]) + cb;  w = sigmoid(A);  out = ys*w + r*(1-w)
//   GSM: out = ys + r
//   cat = [interleave(out) | x[:, fold:]]
// Given d_cat this file produces dx (all C channels, optionally + `add`) and the parameter gradients.
//
// Kernels (every reduction in a fixed order, no atomics):
//   1. gsf_bwd_dwgt   per frame: dA[f,ch] = w(1-w) * sum_p dout*(ys - r)                                   (GSF only)
//   2. gsf_bwd_plane  per frame: dP0/dP1 = transposed 3x3 conv of dA over the (channel, time) plane; per-frame
//                     partials of the channel_conv weight/bias gradients                                    (GSF only)
//   3. gsf_bwd_pix    per (frame, pixel, group): d_r, d_y -> direct part of dx (dxa) and dpre = dgate*(1-gate^2)
//   4. gsf_bwd_z      per element: dz = conv3d^T(dpre);  dbn = dz * (bn(x) > 0)
//   5. BatchNorm backward reduction over dbn (shared column-reduction kernel), then
//      gsf_bwd_final  dx[:, :fold] = dxa + scale*(dbn - mean(dbn) - xhat*mean(dbn*xhat)) (+add);  dx[:, fold:] = d_cat (+add)
//   6. gsf_bwd_w3d    per (frame, row block): dW3d partials = sum_p dpre[f-kt+1, p] * z[f, p + (ky-1, kx-1)]
#include "train_reduce.cuh"

namespace tdeed {

constexpr int GB_THREADS = 256;

__device__ __forceinline__ int gs_out_pos(int ch, int half, int quarter) {
  // interleaved output column of original channel ch:  out[g*half + 2i + k] = in[g*half + k*quarter + i]
  const int g = ch / half, jj = ch - g * half;
  const int k = jj / quarter, i = jj - k * quarter;
  return g * half + 2 * i + k;
}

// ---- 1. dA ----
template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
gsf_bwd_dwgt_kernel(const T* __restrict__ x, const T* __restrict__ dcat, int clip_len, int hw, int c, int fold,
                    const float* __restrict__ gate, const float* __restrict__ wgt, float* __restrict__ dA) {
  extern __shared__ float smem[];
  const int f = blockIdx.x, t = f % clip_len;
  const int half = fold / 2, quarter = fold / 4;
  const int SEG = GB_THREADS / fold > 0 ? GB_THREADS / fold : 1;
  const T* xf = x + (size_t)f * hw * c;
  const T* df = dcat + (size_t)f * hw * c;
  const float* gf = gate + (size_t)f * hw * 2;
  for (int q = threadIdx.x; q < fold * SEG; q += GB_THREADS) {
    const int ch = q % fold, seg = q / fold;
    const int g = ch / half;
    const int jo = gs_out_pos(ch, half, quarter);
    const bool has_src = g == 0 ? (t + 1 < clip_len) : (t > 0);
    const long long soff = g == 0 ? (long long)hw : -(long long)hw;   // frame the shifted y comes from
    float s = 0.f;
    for (int p = seg; p < hw; p += SEG) {
      const float xv = Elem<T>::ld(xf + (size_t)p * c + ch);
      const float r = xv - gf[2 * p + g] * xv;
      float ys = 0.f;
      if (has_src) ys = gf[2 * (p + soff) + g] * Elem<T>::ld(xf + ((long long)p + soff) * c + ch);
      s = fmaf(Elem<T>::ld(df + (size_t)p * c + jo), ys - r, s);
    }
    smem[seg * fold + ch] = s;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < fold; ch += GB_THREADS) {
    float s = 0.f;
    for (int seg = 0; seg < SEG; ++seg) s += smem[seg * fold + ch];
    const float wv = wgt[(size_t)f * fold + ch];
    dA[(size_t)f * fold + ch] = s * wv * (1.f - wv);
  }
}

// ---- 2. plane gradients + channel_conv parameter partials ----
__global__ void __launch_bounds__(GB_THREADS)
gsf_bwd_plane_kernel(const float* __restrict__ dA, const float* __restrict__ sums, int clip_len, int hw, int fold,
                     const float* __restrict__ cc_w, float* __restrict__ dP0, float* __restrict__ dP1, float* __restrict__ ccpart) {
  const int f = blockIdx.x, t = f % clip_len;
  const int half = fold / 2;
  const float inv = 1.f / (float)hw;
  for (int ch = threadIdx.x; ch < fold; ch += GB_THREADS) {
    const int g = ch / half, ci = ch - g * half;
    const float* wk = cc_w + g * 18;
    float a0 = 0.f, a1 = 0.f;
    for (int dc = -1; dc <= 1; ++dc) {
      const int cc = ci - dc;                       // A[cc, tt] used P[cc + dc, tt + dt] = P[ci, t]
      if (cc < 0 || cc >= half) continue;
      for (int dt = -1; dt <= 1; ++dt) {
        const int tt = t - dt;
        if (tt < 0 || tt >= clip_len) continue;
        const float d = dA[((size_t)(f - t + tt)) * fold + g * half + cc];
        a0 = fmaf(wk[(dc + 1) * 3 + (dt + 1)], d, a0);
        a1 = fmaf(wk[9 + (dc + 1) * 3 + (dt + 1)], d, a1);
      }
    }
    dP0[(size_t)f * fold + ch] = a0 * inv;          // already divided by hw: gradient per pixel of ys / r
    dP1[(size_t)f * fold + ch] = a1 * inv;
  }
  // parameter partials of this frame: thread (g, k) with k < 18 weights, k == 18 bias
  if (threadIdx.x < 38) {
    const int g = threadIdx.x / 19, k = threadIdx.x % 19;
    float s = 0.f;
    if (k == 18) {
      for (int ci = 0; ci < half; ++ci) s += dA[(size_t)f * fold + g * half + ci];
    } else {
      const int pl = k / 9, dc = (k % 9) / 3 - 1, dt = k % 3 - 1;
      const int tt = t + dt;
      if (tt >= 0 && tt < clip_len) {
        const int ts = pl == 0 ? (g == 0 ? tt + 1 : tt - 1) : tt;   // plane 0 holds the SHIFTED y
        if (ts >= 0 && ts < clip_len) {
          for (int ci = 0; ci < half; ++ci) {
            const int cc = ci + dc;
            if (cc < 0 || cc >= half) continue;
            const float pv = sums[(((size_t)(f - t + ts)) * fold + g * half + cc) * 2 + pl] * inv;
            s = fmaf(dA[(size_t)f * fold + g * half + ci], pv, s);
          }
        }
      }
    }
    ccpart[(size_t)f * 38 + threadIdx.x] = s;
  }
}

// ---- 3. per (frame, pixel, group) ----
template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
gsf_bwd_pix_kernel(const T* __restrict__ x, const T* __restrict__ dcat, int clip_len, int hw, int c, int fold, int mode,
                   const float* __restrict__ gate, const float* __restrict__ wgt, const float* __restrict__ dP0,
                   const float* __restrict__ dP1, long long total, float* __restrict__ dxa, float* __restrict__ dpre) {
  const long long idx = (long long)blockIdx.x * GB_THREADS + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx & 1);
  const long long fp = idx >> 1;
  const long long f = fp / hw;
  const int t = (int)(f % clip_len);
  const int half = fold / 2, quarter = fold / 4;
  // frame whose output holds y[f] after the shift: group 0: out[t-1] = y[t];  group 1: out[t+1] = y[t]
  const bool has_dst = g == 0 ? (t > 0) : (t + 1 < clip_len);
  const long long doff = g == 0 ? -(long long)hw : (long long)hw;
  const long long fd = g == 0 ? f - 1 : f + 1;
  const T* xt = x + (size_t)fp * c;
  const T* dt_ = dcat + (size_t)fp * c;
  const T* dd = dcat + (size_t)(fp + doff) * c;
  const float gv = gate[(size_t)fp * 2 + g];
  const bool gsf = mode == TDEED_SHIFT_GSF;
  float dgate = 0.f;
  int cq = 0, ck = 0;                              // ci = ck * quarter + cq, tracked without divisions
  for (int ci = 0; ci < half; ++ci) {
    const int ch = g * half + ci;
    const int jo = g * half + 2 * cq + ck;
    if (++cq == quarter) { cq = 0; ++ck; }
    const float dout = Elem<T>::ld(dt_ + jo);
    float d_r, d_ys = 0.f;
    if (gsf) {
      d_r = dout * (1.f - wgt[(size_t)f * fold + ch]) + dP1[(size_t)f * fold + ch];
      if (has_dst) d_ys = Elem<T>::ld(dd + jo) * wgt[(size_t)fd * fold + ch] + dP0[(size_t)fd * fold + ch];
    } else {
      d_r = dout;
      if (has_dst) d_ys = Elem<T>::ld(dd + jo);
    }
    const float d_y = d_ys - d_r;
    const float xv = Elem<T>::ld(xt + ch);
    dgate = fmaf(d_y, xv, dgate);
    dxa[(size_t)fp * fold + ch] = fmaf(d_y, gv, d_r);
  }
  dpre[(size_t)fp * 2 + g] = dgate * (1.f - gv * gv);
}

// ---- 4. conv3d transpose + ReLU mask ----
// Thread = one pixel: its 27 (x 2 gate groups) dpre neighbours are loaded ONCE into registers (bounds handled there), then the
// channel loop is 27 FMAs against weights read as broadcast 16 B shared-memory vectors.  x and dbn pass through shared
// memory in 32-channel chunks so that global accesses stay coalesced.  (One thread per (pixel, channel) re-did the 27 bounds
// checks / address computations / global loads for every channel and was instruction-issue bound, ~8x off the HBM time.)
constexpr int GZ_PIX = 128;      // pixels (threads) per CTA
constexpr int GZ_CH = 32;        // channels per staging chunk
constexpr int GZ_LD = GZ_CH + 1;
static size_t gsf_bwd_z_smem(int fold) { return ((size_t)fold * 28 + 2 * fold + 2 * GZ_PIX * GZ_LD) * sizeof(float); }

__device__ __forceinline__ float dot27(const float* __restrict__ wk, const float (&d)[27]) {
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const float4 wv = *reinterpret_cast<const float4*>(wk + 4 * q);
    acc = fmaf(wv.x, d[4 * q], acc);
    acc = fmaf(wv.y, d[4 * q + 1], acc);
    acc = fmaf(wv.z, d[4 * q + 2], acc);
    acc = fmaf(wv.w, d[4 * q + 3], acc);
  }
  const float4 wv = *reinterpret_cast<const float4*>(wk + 24);
  acc = fmaf(wv.x, d[24], acc);
  acc = fmaf(wv.y, d[25], acc);
  acc = fmaf(wv.z, d[26], acc);
  return acc;
}

template <typename T>
__global__ void __launch_bounds__(GZ_PIX)
gsf_bwd_z_kernel(const T* __restrict__ x, int clip_len, int h, int w, int c, int fold, const float* __restrict__ stats,
                 const float* __restrict__ w3d, const float* __restrict__ dpre, long long M, float* __restrict__ dbn) {
  extern __shared__ __align__(16) float s_gz[];
  float* s_w = s_gz;                        // [fold][28]
  float* s_sc = s_w + (size_t)fold * 28;    // [fold] scale, [fold] shift
  float* s_x = s_sc + 2 * fold;             // [GZ_PIX][GZ_LD]
  float* s_o = s_x + GZ_PIX * GZ_LD;        // [GZ_PIX][GZ_LD]
  const int tid = threadIdx.x;
  for (int i = tid; i < fold * 28; i += GZ_PIX) {
    const int ch = i / 28, tap = i - ch * 28;
    s_w[i] = tap < 27 ? w3d[ch * 27 + tap] : 0.f;
  }
  for (int i = tid; i < 2 * fold; i += GZ_PIX) s_sc[i] = stats[2 * fold + i];
  const int hw = h * w;
  const long long base = (long long)blockIdx.x * GZ_PIX;
  const long long fp = base + tid;
  float d0[27], d1[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) d0[i] = d1[i] = 0.f;
  if (fp < M) {
    const int p = (int)(fp % hw);
    const long long f = fp / hw;
    const int t = (int)(f % clip_len);
    const int py = p / w, px = p - py * w;
    const float2* dp2 = reinterpret_cast<const float2*>(dpre);
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int tt = t - (kt - 1);
      if (tt < 0 || tt >= clip_len) continue;
      const float2* dp = dp2 + (size_t)(f - t + tt) * hw;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = py - (ky - 1);
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = px - (kx - 1);
          if (xx < 0 || xx >= w) continue;
          const float2 v = dp[yy * w + xx];
          d0[kt * 9 + ky * 3 + kx] = v.x;
          d1[kt * 9 + ky * 3 + kx] = v.y;
        }
      }
    }
  }
  const int half = fold / 2;
  const int npx = (int)((M - base) < GZ_PIX ? (M - base) : GZ_PIX);
  for (int c0 = 0; c0 < fold; c0 += GZ_CH) {
    const int nc = min(GZ_CH, fold - c0);
    __syncthreads();                         // previous chunk's s_o readers are done (and s_w / s_sc are filled)
    for (int i = tid; i < npx * nc; i += GZ_PIX) {
      const int pi = i / nc, j = i - pi * nc;
      s_x[pi * GZ_LD + j] = Elem<T>::ld(x + (size_t)(base + pi) * c + c0 + j);
    }
    __syncthreads();
    for (int j = 0; j < nc; ++j) {
      const int ch = c0 + j;                  // uniform over the CTA
      const float xv = s_x[tid * GZ_LD + j];
      const float* wk = s_w + ch * 28;
      const float dz = ch < half ? dot27(wk, d0) : dot27(wk, d1);
      s_o[tid * GZ_LD + j] = fmaf(xv, s_sc[ch], s_sc[fold + ch]) > 0.f ? dz : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < npx * nc; i += GZ_PIX) {
      const int pi = i / nc, j = i - pi * nc;
      dbn[(size_t)(base + pi) * fold + c0 + j] = s_o[pi * GZ_LD + j];
    }
  }
}

// ---- 5. BatchNorm backward over the fold slice ----
template <typename T>
struct GsfBnOp {
  const float* dbn;     // [M, fold]
  const T* x;           // [M, ld]
  long long ld;
  int fold;
  const float* stats;
  float mu[8], is[8];
  static constexpr int kBatch = 1;
  __device__ void begin(int ch0, int nch) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = j < nch ? stats[ch0 + j] : 0.f;
      is[j] = j < nch ? stats[fold + ch0 + j] : 0.f;
    }
  }
  __device__ void row(long long r, int ch0, int nch, float (&a0)[8], float (&a1)[8]) {
    float xv[8];
    load_n(x + r * ld + ch0, nch, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < nch) {
        const float g = dbn[r * fold + ch0 + j];
        a0[j] += g;
        a1[j] = fmaf(g, (xv[j] - mu[j]) * is[j], a1[j]);
      }
    }
  }
};

__global__ void gsf_bn_final_kernel(const float* __restrict__ part, int nparts, long long M, int C, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, float* __restrict__ coef) {
  const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp per channel
  if (ch >= C) return;
  const int cpad = ((C + 7) / 8) * 8;
  double s, q;
  warp_partial_sums(part, nparts, cpad, ch, s, q);
  if ((threadIdx.x & 31) != 0) return;
  dbeta[ch] = (float)s;
  dgamma[ch] = (float)q;
  coef[ch] = (float)(s / (double)M);
  coef[C + ch] = (float)(q / (double)M);
}

template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
gsf_bwd_final_kernel(const T* __restrict__ x, const T* __restrict__ dcat, const T* __restrict__ add, int c, int fold,
                     const float* __restrict__ stats, const float* __restrict__ coef, const float* __restrict__ dxa,
                     const float* __restrict__ dbn, long long total8, T* __restrict__ dx) {
  const long long q = (long long)blockIdx.x * GB_THREADS + threadIdx.x;
  if (q >= total8) return;
  const int c8n = c / 8;
  int c0;
  long long fp;
  if (total8 <= 0x7fffffffLL) {                   // 32-bit index math (a 64-bit division costs ~90 instructions)
    const unsigned q32 = (unsigned)q, fq = q32 / (unsigned)c8n;
    c0 = (int)(q32 - fq * (unsigned)c8n) * 8;
    fp = fq;
  } else {
    c0 = (int)(q % c8n) * 8;
    fp = q / c8n;
  }
  float xv[8], dv[8], av[8], o[8];
  load8(x + q * 8, xv);
  load8(dcat + q * 8, dv);
  if (add) load8(add + q * 8, av);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = c0 + j;
    float v;
    if (ch < fold) {
      const float xhat = (xv[j] - stats[ch]) * stats[fold + ch];
      v = dxa[fp * fold + ch] + stats[2 * fold + ch] * (dbn[fp * fold + ch] - coef[ch] - xhat * coef[fold + ch]);
    } else {
      v = dv[j];
    }
    o[j] = add ? v + av[j] : v;
  }
  store8(dx + q * 8, o);
}

// ---- 6. conv3D weight gradient partials.  grid (row blocks, frames) ----
// Work item = (channel, kt) with the nine (ky, kx) taps in registers: per pixel one dpre value and a sliding 3x3 window of z
// (3 new shared-memory reads per step) feed 9 FMAs — 2.25 FMA per LDS instead of 0.5 with one tap per thread.  When there are
// fewer items than threads (small fold) the rows of the tile are dealt out to several thread slices per item and the slices are
// added in slice order through shared memory.
template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
gsf_bwd_w3d_kernel(const T* __restrict__ x, int clip_len, int h, int w, int c, int fold, int rows_per_cta,
                   const float* __restrict__ stats, const float* __restrict__ dpre, float* __restrict__ part) {
  extern __shared__ float smem[];
  const int wp = w + 2, rp = rows_per_cta + 2;
  const int plane = (rp * wp) | 1;                 // odd channel stride: lanes on different channels hit different banks
  float* s_z = smem;                               // [fold][plane]   rows y0-1 .. y0+rows, cols -1 .. w
  float* s_d = smem + (size_t)fold * ((size_t)rp * wp + 1);        // [3 kt][2 g][rows_per_cta * w]
  float* s_r = s_d + (size_t)6 * rows_per_cta * w;                 // [GB_THREADS][9] slice partials
  const int f = blockIdx.y, t = f % clip_len;
  const int y0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, h - y0);
  const int hw = h * w;
  const T* xf = x + (size_t)f * hw * c;
  for (int i = threadIdx.x; i < (rows + 2) * wp * fold; i += GB_THREADS) {
    const int ch = i % fold, pix = i / fold;
    const int px = pix % wp - 1, py = pix / wp - 1 + y0;
    float v = 0.f;
    if (py >= 0 && py < h && px >= 0 && px < w)
      v = fmaxf(fmaf(Elem<T>::ld(xf + ((size_t)py * w + px) * c + ch), stats[2 * fold + ch], stats[3 * fold + ch]), 0.f);
    s_z[ch * plane + pix] = v;
  }
  const int npx = rows * w;
  for (int i = threadIdx.x; i < 6 * npx; i += GB_THREADS) {
    const int g = i & 1, p = (i >> 1) % npx, kt = i / (2 * npx);
    const int tt = t - (kt - 1);                    // pre[tt] used z[tt + kt - 1] = z[t]
    float v = 0.f;
    if (tt >= 0 && tt < clip_len) v = dpre[((size_t)(f - t + tt) * hw + (size_t)y0 * w + p) * 2 + g];
    s_d[(kt * 2 + g) * npx + p] = v;
  }
  __syncthreads();
  const int half = fold / 2;
  const int items = fold * 3;
  const int nsl = items >= GB_THREADS ? 1 : GB_THREADS / items;
  float* o = part + ((size_t)f * gridDim.x + blockIdx.x) * fold * 27;
  for (int it0 = 0; it0 < items; it0 += GB_THREADS) {
    const int item = nsl == 1 ? it0 + (int)threadIdx.x : (int)threadIdx.x % items;
    const int sl = nsl == 1 ? 0 : (int)threadIdx.x / items;
    const bool active = item < items && sl < nsl;
    const int ch = item / 3, kt = item - ch * 3;
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    if (active) {
      const float* d = s_d + (kt * 2 + ch / half) * npx;
      const float* z = s_z + ch * plane;
      for (int py = sl; py < rows; py += nsl) {
        const float* z0 = z + py * wp;               // z rows y0+py-1, y0+py, y0+py+1 (ky = 0, 1, 2), starting at column -1
        const float* z1 = z0 + wp;
        const float* z2 = z1 + wp;
        const float* dr = d + py * w;
        float a0 = z0[0], a1 = z0[1], b0 = z1[0], b1 = z1[1], c0 = z2[0], c1 = z2[1];
        for (int px = 0; px < w; ++px) {
          const float a2 = z0[px + 2], b2 = z1[px + 2], c2 = z2[px + 2];
          const float dv = dr[px];
          acc[0] = fmaf(dv, a0, acc[0]);
          acc[1] = fmaf(dv, a1, acc[1]);
          acc[2] = fmaf(dv, a2, acc[2]);
          acc[3] = fmaf(dv, b0, acc[3]);
          acc[4] = fmaf(dv, b1, acc[4]);
          acc[5] = fmaf(dv, b2, acc[5]);
          acc[6] = fmaf(dv, c0, acc[6]);
          acc[7] = fmaf(dv, c1, acc[7]);
          acc[8] = fmaf(dv, c2, acc[8]);
          a0 = a1; a1 = a2; b0 = b1; b1 = b2; c0 = c1; c1 = c2;
        }
      }
    }
    if (nsl > 1) {                                  // single pass in this case (items < GB_THREADS)
#pragma unroll
      for (int k = 0; k < 9; ++k) s_r[threadIdx.x * 9 + k] = acc[k];
      __syncthreads();
      if (active && sl == 0)
        for (int q = 1; q < nsl; ++q)
#pragma unroll
          for (int k = 0; k < 9; ++k) acc[k] += s_r[(q * items + item) * 9 + k];
    }
    if (active && sl == 0) {
#pragma unroll
      for (int k = 0; k < 9; ++k) o[ch * 27 + kt * 9 + k] = acc[k];
    }
  }
}

struct GsfBwdPlan {
  int rows, rblocks;
  size_t smem_w3d;
  // workspace offsets (floats)
  size_t dA, dP0, dP1, ccpart, dpre, dxa, dbn, bnpart, coef, w3dpart, colsum, total;
};

static GsfBwdPlan gsf_bwd_plan(int clips, int clip_len, int h, int w, int fold) {
  GsfBwdPlan p;
  const size_t n = (size_t)clips * clip_len, hw = (size_t)h * w;
  int rows = h;
  auto need = [&](int r) { return ((size_t)fold * ((size_t)(r + 2) * (w + 2) + 1) + (size_t)6 * r * w + GB_THREADS * 9) * sizeof(float); };
  // small tiles: 3+ CTAs per SM hide the staging phase of one behind the FMA phase of the others; the price is the halo rows
  while (rows > 4 && need(rows) > 76 * 1024) rows = (rows + 1) / 2;
  while (rows > 1 && need(rows) > 200 * 1024) rows = (rows + 1) / 2;
  p.rows = rows;
  p.rblocks = ceil_div(h, rows);
  p.smem_w3d = need(rows);
  const size_t cpad = ((fold + 7) / 8) * 8;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t at = o; o += (cnt + 3) & ~(size_t)3; return at; };
  p.dA = take(n * fold);
  p.dP0 = take(n * fold);
  p.dP1 = take(n * fold);
  p.ccpart = take(n * 38);
  p.dpre = take(n * hw * 2);
  p.dxa = take(n * hw * fold);
  p.dbn = take(n * hw * fold);
  p.bnpart = take((size_t)BN_MAX_GRID * 2 * cpad);
  p.coef = take(2 * cpad);
  p.w3dpart = take(n * p.rblocks * fold * 27);
  p.colsum = take((size_t)kNumSMs * 2 * 2);
  p.total = o;
  return p;
}

template <typename T>
static int run_gsf_bwd(int mode, const void* x, const void* dcat, const void* add, int clips, int clip_len, int h, int w, int c,
                       int fold, const float* stats, const float* w3d, const float* cc_w, const float* fwd_ws, float* ws,
                       void* dx, float* dw3d, float* db3d, float* dcc, float* dgamma, float* dbeta, cudaStream_t st) {
  const int n = clips * clip_len, hw = h * w;
  TDEED_REQUIRE(n <= 65535, TDEED_ERR_UNSUPPORTED, "tdeed_gsf_bwd: %d frames per call (frames sit on grid.y of the dW3d kernel: at most 65535; split the batch)", n);
  const long long M = (long long)n * hw;
  const float* gate = fwd_ws;
  const float* sums = gate + (size_t)n * hw * 2;
  const float* wgt = sums + (size_t)n * fold * 2;
  const GsfBwdPlan p = gsf_bwd_plan(clips, clip_len, h, w, fold);
  float *dA = ws + p.dA, *dP0 = ws + p.dP0, *dP1 = ws + p.dP1, *ccpart = ws + p.ccpart, *dpre = ws + p.dpre, *dxa = ws + p.dxa,
        *dbn = ws + p.dbn, *bnpart = ws + p.bnpart, *coef = ws + p.coef, *w3dpart = ws + p.w3dpart, *cs = ws + p.colsum;
  int rc;
  if (mode == TDEED_SHIFT_GSF) {
    const int SEG = GB_THREADS / fold > 0 ? GB_THREADS / fold : 1;
    gsf_bwd_dwgt_kernel<T><<<n, GB_THREADS, (size_t)SEG * fold * sizeof(float), st>>>((const T*)x, (const T*)dcat, clip_len, hw, c, fold, gate, wgt, dA);
    if ((rc = check_launch("tdeed_gsf_bwd(dwgt)"))) return rc;
    gsf_bwd_plane_kernel<<<n, GB_THREADS, 0, st>>>(dA, sums, clip_len, hw, fold, cc_w, dP0, dP1, ccpart);
    if ((rc = check_launch("tdeed_gsf_bwd(plane)"))) return rc;
    launch_partial_sum(ccpart, n, 38, dcc, st);
    if ((rc = check_launch("tdeed_gsf_bwd(cc)"))) return rc;
  }
  gsf_bwd_pix_kernel<T><<<(unsigned)ceil_div_ll(M * 2, GB_THREADS), GB_THREADS, 0, st>>>(
      (const T*)x, (const T*)dcat, clip_len, hw, c, fold, mode, gate, wgt, dP0, dP1, M * 2, dxa, dpre);
  if ((rc = check_launch("tdeed_gsf_bwd(pix)"))) return rc;
  {
    auto kz = gsf_bwd_z_kernel<T>;
    const size_t smem_z = gsf_bwd_z_smem(fold);
    static size_t z_set = 48 * 1024;
    if (smem_z > z_set) {
      cudaError_t e = cudaFuncSetAttribute(kz, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_gsf_bwd: cudaFuncSetAttribute(z): %s", cudaGetErrorString(e));
      z_set = 200 * 1024;
    }
    TDEED_REQUIRE(smem_z <= 200 * 1024, TDEED_ERR_UNSUPPORTED, "tdeed_gsf_bwd: fold %d too large for the conv3d^T kernel", fold);
    kz<<<(unsigned)ceil_div_ll(M, GZ_PIX), GZ_PIX, smem_z, st>>>((const T*)x, clip_len, h, w, c, fold, stats, w3d, dpre, M, dbn);
  }
  if ((rc = check_launch("tdeed_gsf_bwd(z)"))) return rc;
  GsfBnOp<T> op;
  op.dbn = dbn;
  op.x = (const T*)x;
  op.ld = c;
  op.fold = fold;
  op.stats = stats;
  const int grid = bn_grid(M, fold);
  bn_reduce_kernel<T, GsfBnOp<T>><<<grid, BN_THREADS, 0, st>>>(op, M, fold, bnpart);
  if ((rc = check_launch("tdeed_gsf_bwd(bn partial)"))) return rc;
  gsf_bn_final_kernel<<<ceil_div(fold, 8), 256, 0, st>>>(bnpart, grid, M, fold, dgamma, dbeta, coef);
  if ((rc = check_launch("tdeed_gsf_bwd(bn final)"))) return rc;
  const long long total8 = M * (c / 8);
  gsf_bwd_final_kernel<T><<<(unsigned)ceil_div_ll(total8, GB_THREADS), GB_THREADS, 0, st>>>(
      (const T*)x, (const T*)dcat, (const T*)add, c, fold, stats, coef, dxa, dbn, total8, (T*)dx);
  if ((rc = check_launch("tdeed_gsf_bwd(final)"))) return rc;
  auto kw = gsf_bwd_w3d_kernel<T>;
  static size_t w_set = 48 * 1024;
  if (p.smem_w3d > w_set) {
    cudaError_t e = cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_gsf_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    w_set = 200 * 1024;
  }
  TDEED_REQUIRE(p.smem_w3d <= 200 * 1024, TDEED_ERR_UNSUPPORTED, "tdeed_gsf_bwd: a row of %d px x %d ch does not fit shared memory", w, fold);
  kw<<<dim3(p.rblocks, n), GB_THREADS, p.smem_w3d, st>>>((const T*)x, clip_len, h, w, c, fold, p.rows, stats, dpre, w3dpart);
  if ((rc = check_launch("tdeed_gsf_bwd(w3d)"))) return rc;
  launch_partial_sum(w3dpart, n * p.rblocks, (long long)fold * 27, dw3d, st);
  if ((rc = check_launch("tdeed_gsf_bwd(w3d final)"))) return rc;
  return tdeed_colsum(TDEED_F32, dpre, M, 2, 2, db3d, cs, st);
}

}  // namespace tdeed

extern "C" long long tdeed_gsf_bwd_workspace_floats(int clips, int clip_len, int h, int w, int fold) {
  return (long long)tdeed::gsf_bwd_plan(clips, clip_len, h, w, fold).total;
}

extern "C" int tdeed_gsf_bwd(int dtype, int mode, const void* x, const void* dcat, const void* add, int clips, int clip_len,
                             int h, int w, int c, int fold, const float* bn_stats, const float* conv3d_w, const float* cc_w,
                             const float* fwd_workspace, float* workspace, void* dx, float* d_conv3d_w, float* d_conv3d_b,
                             float* d_cc, float* d_bn_gamma, float* d_bn_beta, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && dcat && bn_stats && conv3d_w && fwd_workspace && workspace && dx && d_conv3d_w && d_conv3d_b && d_bn_gamma &&
                d_bn_beta, TDEED_ERR_SHAPE, "tdeed_gsf_bwd: null pointer");
  TDEED_REQUIRE(mode == TDEED_SHIFT_GSM || (cc_w && d_cc), TDEED_ERR_SHAPE, "tdeed_gsf_bwd: GSF needs channel_conv weights");
  TDEED_REQUIRE(clips > 0 && clip_len > 0 && h > 0 && w > 0 && fold > 0 && fold % 4 == 0 && fold <= c && fold <= 1024 && c % 8 == 0 &&
                (long long)clips * clip_len <= 65535, TDEED_ERR_SHAPE, "tdeed_gsf_bwd: bad shape clips=%d T=%d %dx%dx%d fold=%d",
                clips, clip_len, h, w, c, fold);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16)
    return run_gsf_bwd<__nv_bfloat16>(mode, x, dcat, add, clips, clip_len, h, w, c, fold, bn_stats, conv3d_w, cc_w, fwd_workspace,
                                      workspace, dx, d_conv3d_w, d_conv3d_b, d_cc, d_bn_gamma, d_bn_beta, st);
  if (dtype == TDEED_F32)
    return run_gsf_bwd<float>(mode, x, dcat, add, clips, clip_len, h, w, c, fold, bn_stats, conv3d_w, cc_w, fwd_workspace, workspace,
                              dx, d_conv3d_w, d_conv3d_b, d_cc, d_bn_gamma, d_bn_beta, st);
  set_error("tdeed_gsf_bwd: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}
