// (5) Gate-Shift (GSM) / Gate-Shift-Fuse (GSF) on the first `fold` channels of an NHWC activation.
// Reference: model/shift.py:64-93, model/impl/gsm.py:89-116, model/impl/gsf.py:38-93 (eval-mode BN3d).
//
//   gate[b,g,t,p] = tanh( conv3d_{3x3x3, pad 1}( relu(bn(x)) )[group g] )
//   y = gate * x, r = x - y;  spatial sums of y and r per (frame, channel)
//   GSF: w[b,c,t] = sigmoid( conv2d_{2->1,3x3,pad1} over the (channel,time) plane of [mean(shift(y)), mean(r)] )
//   out = shift(y)*w + r*(1-w)   (GSM: shift(y) + r), channel-interleaved
//   shift: group 0 takes y from t+1 (zero at T-1), group 1 from t-1 (zero at 0); no leakage across clips.
//
// Kernel plan (every input frame is read from HBM once per kernel, all reductions in a fixed order):
//   1. gsf_q_kernel     per (frame, row block): z = relu(bn(x)) staged channel-major in shared memory, then the
//                       three temporal slices of the 3x3x3 kernel are applied as 2D convs:
//                       Q[f][p][3g+kt] = (W3d[g][:, kt] * z[f])[p].  The temporal sum is deferred, so no CTA
//                       needs its neighbour frames.
//   2. gsf_gate_kernel  per frame: gate[t] = tanh(b + Q[t-1][kt=0] + Q[t][kt=1] + Q[t+1][kt=2]) and the
//                       per-channel spatial sums of y and r.
//   3. gsf_weight_kernel per frame (GSF only): the 3x3 fusion conv over the (channel, time) plane + sigmoid.
//   4. gsf_blend_kernel  elementwise blend + channel interleave, 8 output channels (16 B) per thread; pad columns
//                       (ld_out > fold) are written as zeros because the GEMM multiplies them by zero weights.
// The output only holds the `fold` channels; the following 1x1 conv reads it as the first K-segment of a virtual
// concat with the untouched channels of x (tdeed_gemm_fwd segments).
//
// workspace layout (floats): gate [N*hw*2] | sums [N*fold*2] (y, r) | wgt [N*fold] | Q [N*hw*6]
#include "common.cuh"

namespace tdeed {

constexpr int GS_THREADS = 256;
constexpr int GS_Q_SMEM_BUDGET = 96 * 1024;

// ---- kernel 1: Q maps.  grid (row_blocks, frames) ----
template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gsf_q_kernel(const T* __restrict__ x, int h, int w, int c, int fold, int rows_per_cta, int nsl,
             const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
             const float* __restrict__ w3d, float* __restrict__ Q) {
  extern __shared__ __align__(16) float smem[];
  const int half = fold / 2;
  const int wp = w + 2;
  const int rp = rows_per_cta + 2;
  const int plane = (rp * wp + 4) | 1;           // +4: the last strip may read past the row end; odd pitch spreads planes over banks
  float4* s_w = reinterpret_cast<float4*>(smem); // [9][fold] : (kt0, kt1, kt2, -)
  float* s_z = smem + 9 * fold * 4;              // [fold][plane]
  const int f = blockIdx.y;
  const int y0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, h - y0);

  for (int i = threadIdx.x; i < 9 * fold; i += GS_THREADS) {
    const int ch = i % fold, tap = i / fold;
    const float* wsrc = w3d + (size_t)ch * 27 + tap;      // [2][half][3][3][3] flattened: ch = g*half + ci
    s_w[i] = make_float4(wsrc[0], wsrc[9], wsrc[18], 0.f);
  }
  // stage z = relu(bn(x)) for rows [y0-1, y0+rows] x cols [-1, w], zero outside the image
  const T* xf = x + (size_t)f * h * w * c;
  const int c8n = fold / 8, ctail = fold - c8n * 8;
  for (int i = threadIdx.x; i < (rows + 2) * wp * (c8n + (ctail ? 1 : 0)); i += GS_THREADS) {
    const int cg = i % (c8n + (ctail ? 1 : 0));
    const int pix = i / (c8n + (ctail ? 1 : 0));
    const int px = pix % wp - 1, py = pix / wp - 1 + y0;
    const int ch0 = cg * 8;
    const int nch = (cg < c8n) ? 8 : ctail;
    float v[8];
    const bool inside = py >= 0 && py < h && px >= 0 && px < w;
    if (inside) {
      const T* src = xf + ((size_t)py * w + px) * c + ch0;
      if (nch == 8) {
        load8(src, v);
      } else {
        for (int j = 0; j < nch; ++j) v[j] = Elem<T>::ld(src + j);
      }
    }
    for (int j = 0; j < nch; ++j)
      s_z[(ch0 + j) * plane + pix] = inside ? fmaxf(fmaf(v[j], bn_scale[ch0 + j], bn_shift[ch0 + j]), 0.f) : 0.f;
  }
  __syncthreads();

  // compute: a thread owns a strip of 4 horizontally adjacent pixels, so every z value feeds up to 3 taps x 3 temporal
  // slices and every weight fetch feeds 4 pixels (the one-pixel-per-thread version ran at 85 % L1TEX: shared-memory bound).
  // Small frames have few strips (14 at 7x7, 56 at 14x14): the channels of each group are then dealt out to `nsl` thread
  // slices per strip and the slices are added through shared memory in slice order (with one slice only 14 of the 256
  // threads worked at 7x7: 232 us per launch for an 81 MB problem).
  const int hw = h * w;
  const int strips = (w + 3) / 4;
  const int nstr = rows * strips;
  const int cps = (half + nsl - 1) / nsl;           // channels of a group per slice
  float* s_red = s_z;                               // reused after the barrier below (host checks that it fits)
  for (int it0 = 0; it0 < nstr * nsl; it0 += GS_THREADS) {
    const int it = it0 + threadIdx.x;
    const bool active = it < nstr * nsl;
    const int sl = active ? it / nstr : 0;
    const int st = active ? it - sl * nstr : 0;
    const int py = st / strips, x0 = (st - py * strips) * 4;
    float acc[2][3][4];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[g][k][q] = 0.f;
    if (active) {
      const int ci0 = sl * cps, ci1 = min(half, ci0 + cps);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        for (int ci = ci0; ci < ci1; ++ci) {
          const int ch = g * half + ci;
          const float* zp = s_z + ch * plane + py * wp + x0;   // top-left of the strip's 3 x 6 window
          float zr[3][6];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int j = 0; j < 6; ++j) zr[dy][j] = zp[dy * wp + j];   // columns past the row end are only used by masked pixels
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float4 w4 = s_w[(dy * 3 + dx) * fold + ch];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float zv = zr[dy][q + dx];
                acc[g][0][q] = fmaf(zv, w4.x, acc[g][0][q]);
                acc[g][1][q] = fmaf(zv, w4.y, acc[g][1][q]);
                acc[g][2][q] = fmaf(zv, w4.z, acc[g][2][q]);
              }
            }
        }
      }
    }
    if (nsl == 1) {
      if (!active) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (x0 + q >= w) break;
        float2* qo = reinterpret_cast<float2*>(Q + ((size_t)f * hw + (size_t)(y0 + py) * w + x0 + q) * 6);
        qo[0] = make_float2(acc[0][0][q], acc[0][1][q]);
        qo[1] = make_float2(acc[0][2][q], acc[1][0][q]);
        qo[2] = make_float2(acc[1][1][q], acc[1][2][q]);
      }
    } else {
      // nsl > 1 only when nstr * nsl <= GS_THREADS: a single pass, so the barriers below are reached by every thread
      __syncthreads();                              // all z reads are done: the tile becomes the slice-partials buffer
      if (active) {
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) s_red[(sl * nstr + st) * 24 + (g * 3 + k) * 4 + q] = acc[g][k][q];
      }
      __syncthreads();
      for (int o = threadIdx.x; o < nstr * 4; o += GS_THREADS) {
        const int st2 = o >> 2, q = o & 3;
        const int py2 = st2 / strips, x2 = (st2 - py2 * strips) * 4 + q;
        if (x2 >= w) continue;
        float v[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = 0.f;
        for (int s2 = 0; s2 < nsl; ++s2)
#pragma unroll
          for (int j = 0; j < 6; ++j) v[j] += s_red[(s2 * nstr + st2) * 24 + j * 4 + q];
        float2* qo = reinterpret_cast<float2*>(Q + ((size_t)f * hw + (size_t)(y0 + py2) * w + x2) * 6);
        qo[0] = make_float2(v[0], v[1]);
        qo[1] = make_float2(v[2], v[3]);
        qo[2] = make_float2(v[4], v[5]);
      }
    }
  }
}

// ---- kernel 2: gate + per-(frame,channel) sums.  One CTA per frame. ----
template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gsf_gate_kernel(const T* __restrict__ x, int clip_len, int hw, int c, int fold, const float* __restrict__ b3d,
                const float* __restrict__ Q, float* __restrict__ gate, float* __restrict__ sums) {
  extern __shared__ float smem[];
  const int f = blockIdx.x;
  const int t = f % clip_len;
  const int half = fold / 2;
  const float bias0 = b3d[0], bias1 = b3d[1];
  float* gate_f = gate + (size_t)f * hw * 2;
  const float* q0 = Q + (size_t)(f - 1) * hw * 6;
  const float* q1 = Q + (size_t)f * hw * 6;
  const float* q2 = Q + (size_t)(f + 1) * hw * 6;
  for (int p = threadIdx.x; p < hw; p += GS_THREADS) {
    float a0 = bias0 + q1[p * 6 + 1], a1 = bias1 + q1[p * 6 + 4];
    if (t > 0) { a0 += q0[p * 6 + 0]; a1 += q0[p * 6 + 3]; }
    if (t < clip_len - 1) { a0 += q2[p * 6 + 2]; a1 += q2[p * 6 + 5]; }
    gate_f[2 * p] = tanhf(a0);
    gate_f[2 * p + 1] = tanhf(a1);
  }
  __syncthreads();
  // spatial sums of y = gate*x and r = x - y per channel (fixed order -> deterministic)
  const int SEG = GS_THREADS / fold > 0 ? GS_THREADS / fold : 1;
  float* s_part = smem;                 // [SEG][fold][2]
  const T* xf = x + (size_t)f * hw * c;
  for (int q = threadIdx.x; q < fold * SEG; q += GS_THREADS) {
    const int ch = q % fold, seg = q / fold;
    const int g = ch / half;
    float sy = 0.f, sr = 0.f;
#pragma unroll 4
    for (int p = seg; p < hw; p += SEG) {
      const float xv = Elem<T>::ld(xf + (size_t)p * c + ch);
      const float yv = gate_f[2 * p + g] * xv;
      sy += yv;
      sr += xv - yv;
    }
    s_part[(seg * fold + ch) * 2] = sy;
    s_part[(seg * fold + ch) * 2 + 1] = sr;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < fold; ch += GS_THREADS) {
    float sy = 0.f, sr = 0.f;
    for (int seg = 0; seg < SEG; ++seg) {
      sy += s_part[(seg * fold + ch) * 2];
      sr += s_part[(seg * fold + ch) * 2 + 1];
    }
    sums[((size_t)f * fold + ch) * 2] = sy;
    sums[((size_t)f * fold + ch) * 2 + 1] = sr;
  }
}

// ---- kernel 3: GSF fusion weights.  One CTA per frame, one thread per channel. ----
__global__ void __launch_bounds__(GS_THREADS)
gsf_weight_kernel(const float* __restrict__ sums, int clip_len, int hw, int fold, const float* __restrict__ cc_w,
                  const float* __restrict__ cc_b, float* __restrict__ wgt) {
  const int f = blockIdx.x;
  const int t = f % clip_len;
  const int half = fold / 2;
  const float inv = 1.f / (float)hw;
  for (int ch = threadIdx.x; ch < fold; ch += GS_THREADS) {
    const int g = ch / half, ci = ch - g * half;
    const float* wk = cc_w + g * 18;          // [2 (y, r)][3 (channel)][3 (time)]
    float a = cc_b[g];
    for (int dc = -1; dc <= 1; ++dc) {
      const int cc = ci + dc;
      if (cc < 0 || cc >= half) continue;
      for (int dt = -1; dt <= 1; ++dt) {
        const int tt = t + dt;
        if (tt < 0 || tt >= clip_len) continue;
        // plane 0: mean of the SHIFTED y at time tt (= y at tt+1 for g=0, tt-1 for g=1, zero outside the clip)
        const int ts = (g == 0) ? tt + 1 : tt - 1;
        float ym = 0.f;
        if (ts >= 0 && ts < clip_len) ym = sums[(((size_t)(f - t + ts)) * fold + g * half + cc) * 2] * inv;
        const float rm = sums[(((size_t)(f - t + tt)) * fold + g * half + cc) * 2 + 1] * inv;
        a = fmaf(wk[(dc + 1) * 3 + (dt + 1)], ym, a);
        a = fmaf(wk[9 + (dc + 1) * 3 + (dt + 1)], rm, a);
      }
    }
    wgt[(size_t)f * fold + ch] = sigmoidf_(a);
  }
}

// ---- kernel 4: blend + channel interleave.  Thread per (pixel, 8 output channels). ----
template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gsf_blend_kernel(const T* __restrict__ x, int clip_len, int hw, int c, int fold, int mode,
                 const float* __restrict__ gate, const float* __restrict__ wgt, T* __restrict__ out, int ld_out,
                 long long total, int copy_tail) {
  // grid (frames, pixel-octet blocks): the frame index is the block's, one 32-bit division per thread (the former flat index
  // cost four 64-bit divisions per thread plus one division per output channel: more than half of the kernel's instructions)
  const int o8n = ld_out / 8;
  const int local = blockIdx.y * GS_THREADS + threadIdx.x;   // pixel * o8n + octet inside the frame
  if (local >= hw * o8n) return;
  const int half = fold / 2, quarter = fold / 4;
  const int p = local / o8n;
  const int o8 = local - p * o8n;
  const long long f = blockIdx.x;
  const long long fp = f * hw + p;                  // frame*hw + pixel
  const int t = (int)(blockIdx.x % (unsigned)clip_len);
  const T* xt = x + (size_t)fp * c;
  const float g0 = gate[(size_t)fp * 2], g1 = gate[(size_t)fp * 2 + 1];
  // shifted sources: group 0 <- t+1, group 1 <- t-1 (zero outside the clip)
  const bool has_next = t + 1 < clip_len, has_prev = t > 0;
  const T* xn = xt + (size_t)hw * c;
  const T* xp = xt - (size_t)hw * c;
  const float gn = has_next ? gate[((size_t)fp + hw) * 2] : 0.f;
  const float gp = has_prev ? gate[((size_t)fp - hw) * 2 + 1] : 0.f;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int jo = o8 * 8 + j;                      // output (interleaved) channel, or a pad column
    float v = 0.f;
    if (jo < fold) {
      const int g = jo >= half ? 1 : 0, jj = jo - g * half;       // jo < fold = 2 * half
      const int ch = g * half + (jj & 1) * quarter + (jj >> 1);   // out[2i+k] = in[k*quarter + i]
      const float xv = Elem<T>::ld(xt + ch);
      const float r = xv - (g == 0 ? g0 : g1) * xv;
      float ys = 0.f;
      if (g == 0) { if (has_next) ys = gn * Elem<T>::ld(xn + ch); }
      else { if (has_prev) ys = gp * Elem<T>::ld(xp + ch); }
      if (mode == TDEED_SHIFT_GSF) {
        const float wv = wgt[(size_t)f * fold + ch];
        v = ys * wv + r * (1.f - wv);
      } else {
        v = ys + r;
      }
    }
    else if (copy_tail && jo < c) v = Elem<T>::ld(xt + jo);   // training: materialise the concat [gs(x[:, :fold]) | x[:, fold:]]
    o[j] = v;
  }
  store8(out + (size_t)fp * ld_out + o8 * 8, o);
}

template <typename T>
static int launch_gsf(int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                      const float* bn_scale, const float* bn_shift, const float* w3d, const float* b3d,
                      const float* cc_w, const float* cc_b, float* ws, void* out, int ld_out, int copy_tail, cudaStream_t st) {
  const int n = clips * clip_len, hw = h * w;
  TDEED_REQUIRE(n <= 65535, TDEED_ERR_UNSUPPORTED, "tdeed_gsf_fwd: %d frames per call (the Q-map kernel puts frames on grid.y: at most 65535; split the batch)", n);
  float* gate = ws;
  float* sums = gate + (size_t)n * hw * 2;
  float* wgt = sums + (size_t)n * fold * 2;
  float* Q = wgt + (size_t)n * fold;
  Q += (4 - ((Q - ws) & 3)) & 3;                       // 16-byte align (float2 stores need 8)

  // kernel 1: rows per CTA so that the staged z tile fits the shared-memory budget
  const size_t w_bytes = (size_t)9 * fold * 16;
  int rows = h;
  while (rows > 1 && w_bytes + (size_t)fold * ((size_t)(rows + 2) * (w + 2) + 5) * 4 > (size_t)GS_Q_SMEM_BUDGET) rows = (rows + 1) / 2;
  const size_t smem_q = w_bytes + (size_t)fold * ((size_t)(rows + 2) * (w + 2) + 5) * 4;
  TDEED_REQUIRE(smem_q <= 200 * 1024, TDEED_ERR_UNSUPPORTED, "tdeed_gsf_fwd: a single row of %d px x %d ch does not fit shared memory", w, fold);
  auto kq = gsf_q_kernel<T>;
  static size_t q_set = 48 * 1024;
  if (smem_q > q_set) {
    cudaError_t e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_gsf_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    q_set = 200 * 1024;
  }
  // channel slices per strip for small frames (single compute pass, partials must fit the z tile they replace)
  int nsl = 1;
  {
    const int nstr = (rows < h ? rows : h) * ((w + 3) / 4);
    while (nsl < 16 && nstr * (nsl * 2) <= GS_THREADS && nsl * 2 <= fold / 2 &&
           (size_t)nstr * (nsl * 2) * 24 <= (size_t)fold * ((size_t)(rows + 2) * (w + 2) + 5))
      nsl *= 2;
  }
  kq<<<dim3(ceil_div(h, rows), n), GS_THREADS, smem_q, st>>>((const T*)x, h, w, c, fold, rows, nsl, bn_scale, bn_shift, w3d, Q);
  int rc = check_launch("tdeed_gsf_fwd(q)");
  if (rc) return rc;

  const int SEG = GS_THREADS / fold > 0 ? GS_THREADS / fold : 1;
  gsf_gate_kernel<T><<<n, GS_THREADS, (size_t)SEG * fold * 2 * sizeof(float), st>>>((const T*)x, clip_len, hw, c, fold, b3d, Q, gate, sums);
  rc = check_launch("tdeed_gsf_fwd(gate)");
  if (rc) return rc;
  if (mode == TDEED_SHIFT_GSF) {
    gsf_weight_kernel<<<n, fold < GS_THREADS ? ((fold + 31) / 32 * 32) : GS_THREADS, 0, st>>>(sums, clip_len, hw, fold, cc_w, cc_b, wgt);
    rc = check_launch("tdeed_gsf_fwd(weights)");
    if (rc) return rc;
  }
  const long long total = (long long)n * hw * (ld_out / 8);
  gsf_blend_kernel<T><<<dim3((unsigned)n, (unsigned)ceil_div(hw * (ld_out / 8), GS_THREADS)), GS_THREADS, 0, st>>>(
      (const T*)x, clip_len, hw, c, fold, mode, gate, wgt, (T*)out, ld_out, total, copy_tail);
  return check_launch("tdeed_gsf_fwd(blend)");
}

}  // namespace tdeed

extern "C" long long tdeed_gsf_workspace_floats(int clips, int clip_len, int h, int w, int fold) {
  const long long n = (long long)clips * clip_len;
  return n * h * w * 2 + n * fold * 2 + n * fold + 4 + n * h * w * 6;
}

static int gsf_dispatch(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                        const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                        const float* cc_w, const float* cc_b, float* workspace, void* out, int ld_out, int copy_tail,
                        void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && bn_scale && bn_shift && conv3d_w && conv3d_b && workspace && out, TDEED_ERR_SHAPE,
                "tdeed_gsf_fwd: null pointer");
  TDEED_REQUIRE(mode == TDEED_SHIFT_GSM || (cc_w && cc_b), TDEED_ERR_SHAPE, "tdeed_gsf_fwd: GSF needs channel_conv weights");
  TDEED_REQUIRE(clips > 0 && clip_len > 0 && h > 0 && w > 0 && fold > 0 && fold % 4 == 0 && fold <= c && fold <= 1024 &&
                ld_out >= fold && ld_out % 8 == 0 && c % 8 == 0 && (long long)clips * clip_len <= 65535, TDEED_ERR_SHAPE,
                "tdeed_gsf_fwd: bad shape clips=%d T=%d %dx%dx%d fold=%d ld_out=%d", clips, clip_len, h, w, c, fold, ld_out);
  TDEED_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, TDEED_ERR_SHAPE, "tdeed_gsf_fwd: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16)
    return launch_gsf<__nv_bfloat16>(mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b,
                                     cc_w, cc_b, workspace, out, ld_out, copy_tail, st);
  if (dtype == TDEED_F32)
    return launch_gsf<float>(mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                             workspace, out, ld_out, copy_tail, st);
  set_error("tdeed_gsf_fwd: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}

extern "C" int tdeed_gsf_fwd(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                             const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                             const float* cc_w, const float* cc_b, float* workspace, void* out, int ld_out,
                             void* stream) {
  return gsf_dispatch(dtype, mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                      workspace, out, ld_out, 0, stream);
}

// training: writes the full concat y = [gs(x[:, :fold]) | x[:, fold:]] as [frames*h*w, c] (model/shift.py:89-93); the
// workspace (gate, per-frame sums, fusion weights) is what tdeed_gsf_bwd reads back.
extern "C" int tdeed_gsf_cat_fwd(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                                 const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                                 const float* cc_w, const float* cc_b, float* workspace, void* out, void* stream) {
  return gsf_dispatch(dtype, mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                      workspace, out, c, 1, stream);
}
