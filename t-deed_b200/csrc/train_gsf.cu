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
  for (int ci = 0; ci < half; ++ci) {
    const int ch = g * half + ci;
    const int jo = g * half + 2 * (ci % quarter) + ci / quarter;
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
template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
gsf_bwd_z_kernel(const T* __restrict__ x, int clip_len, int h, int w, int c, int fold, const float* __restrict__ stats,
                 const float* __restrict__ w3d, const float* __restrict__ dpre, long long total, float* __restrict__ dbn) {
  extern __shared__ float s_w[];        // [fold][27]
  for (int i = threadIdx.x; i < fold * 27; i += GB_THREADS) s_w[i] = w3d[i];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * GB_THREADS + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % fold);
  const long long fp = idx / fold;
  const int hw = h * w;
  const int p = (int)(fp % hw);
  const long long f = fp / hw;
  const int t = (int)(f % clip_len);
  const int py = p / w, px = p - py * w;
  const int g = ch / (fold / 2);
  const float xv = Elem<T>::ld(x + (size_t)fp * c + ch);
  float dz = 0.f;
  if (fmaf(xv, stats[2 * fold + ch], stats[3 * fold + ch]) > 0.f) {
    const float* wk = s_w + ch * 27;
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int tt = t - (kt - 1);
      if (tt < 0 || tt >= clip_len) continue;
      const float* dp = dpre + (size_t)(f - t + tt) * hw * 2;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = py - (ky - 1);
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = px - (kx - 1);
          if (xx < 0 || xx >= w) continue;
          dz = fmaf(wk[kt * 9 + ky * 3 + kx], dp[((size_t)yy * w + xx) * 2 + g], dz);
        }
      }
    }
  }
  dbn[idx] = dz;
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
  const int c0 = (int)(q % c8n) * 8;
  const long long fp = q / c8n;
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
template <typename T>
__global__ void __launch_bounds__(GB_THREADS)
gsf_bwd_w3d_kernel(const T* __restrict__ x, int clip_len, int h, int w, int c, int fold, int rows_per_cta,
                   const float* __restrict__ stats, const float* __restrict__ dpre, float* __restrict__ part) {
  extern __shared__ float smem[];
  const int wp = w + 2, rp = rows_per_cta + 2;
  const int plane = rp * wp + 1;
  float* s_z = smem;                               // [fold][plane]   rows y0-1 .. y0+rows, cols -1 .. w
  float* s_d = smem + (size_t)fold * plane;        // [3 kt][2 g][rows_per_cta * w]
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
  float* o = part + ((size_t)f * gridDim.x + blockIdx.x) * fold * 27;
  for (int item = threadIdx.x; item < fold * 27; item += GB_THREADS) {
    const int ch = item / 27, tap = item - ch * 27;
    const int kt = tap / 9, ky = (tap % 9) / 3, kx = tap % 3;
    const int g = ch / half;
    const float* d = s_d + (kt * 2 + g) * npx;
    const float* z = s_z + ch * plane;
    float s = 0.f;
    for (int py = 0; py < rows; ++py) {
      const float* zr = z + (py + ky) * wp + kx;     // z at (y0 + py + ky - 1, px + kx - 1)
      const float* dr = d + py * w;
      for (int px = 0; px < w; ++px) s = fmaf(dr[px], zr[px], s);
    }
    o[item] = s;
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
  auto need = [&](int r) { return ((size_t)fold * ((size_t)(r + 2) * (w + 2) + 1) + (size_t)6 * r * w) * sizeof(float); };
  while (rows > 1 && need(rows) > 160 * 1024) rows = (rows + 1) / 2;
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
    partial_sum_kernel<<<1, 64, 0, st>>>(ccpart, n, 38, dcc);
    if ((rc = check_launch("tdeed_gsf_bwd(cc)"))) return rc;
  }
  gsf_bwd_pix_kernel<T><<<(unsigned)ceil_div_ll(M * 2, GB_THREADS), GB_THREADS, 0, st>>>(
      (const T*)x, (const T*)dcat, clip_len, hw, c, fold, mode, gate, wgt, dP0, dP1, M * 2, dxa, dpre);
  if ((rc = check_launch("tdeed_gsf_bwd(pix)"))) return rc;
  gsf_bwd_z_kernel<T><<<(unsigned)ceil_div_ll(M * fold, GB_THREADS), GB_THREADS, (size_t)fold * 27 * sizeof(float), st>>>(
      (const T*)x, clip_len, h, w, c, fold, stats, w3d, dpre, M * fold, dbn);
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
  partial_sum_kernel<<<ceil_div(fold * 27, 256), 256, 0, st>>>(w3dpart, n * p.rblocks, (long long)fold * 27, dw3d);
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
