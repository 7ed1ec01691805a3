// (5) Gate-Shift (GSM) / Gate-Shift-Fuse (GSF) on the first `fold` channels of an NHWC activation.
// Reference: model/shift.py:64-93, model/impl/gsm.py:89-116, model/impl/gsf.py:38-93 (eval-mode BN3d).
//
//   gate[b,g,t,p] = tanh( conv3d_{3x3x3, pad 1}( relu(bn(x)) )[group g] )          kernel 1 (gate)
//   y = gate * x, r = x - y;  spatial sums of y and r per (frame, channel)           kernel 1 (sums)
//   GSF: w[b,c,t] = sigmoid( conv2d_{2->1,3x3,pad1} over the (channel,time) plane of
//                            [mean(shift(y)), mean(r)] )                              kernel 2 (fuse weights)
//   out = shift(y)*w + r*(1-w)   (GSM: shift(y) + r), channel-interleaved            kernel 3 (blend)
//   shift: group 0 takes y from t+1 (zero at T-1), group 1 from t-1 (zero at 0); no leakage across clips.
// The output only holds the `fold` channels; the following 1x1 conv reads it as the first K-segment
// of a virtual concat with the untouched channels of x (tdeed_gemm_fwd segments).
//
// workspace layout (floats): gate [N*hw*2] | sums [N*fold*2] (y, r) | wgt [N*fold]
#include "common.cuh"

namespace tdeed {

constexpr int GS_THREADS = 256;

// ---- kernel 1: gate + per-(frame,channel) sums.  One CTA per frame. ----
template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gsf_gate_kernel(const T* __restrict__ x, int clip_len, int h, int w, int c, int fold,
                const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                const float* __restrict__ w3d, const float* __restrict__ b3d,
                float* __restrict__ gate, float* __restrict__ sums) {
  extern __shared__ float smem[];
  const int half = fold / 2;
  float* s_w = smem;                   // [27][fold]  (tap-major, channel = g*half + ci)
  float* s_scale = s_w + 27 * fold;    // [fold]
  float* s_shift = s_scale + fold;     // [fold]
  float* s_part = s_shift + fold;      // [SEG][fold][2]
  const int f = blockIdx.x;
  const int t = f % clip_len;
  const int hw = h * w;

  for (int i = threadIdx.x; i < 27 * fold; i += GS_THREADS) {
    const int ch = i % fold, tap = i / fold;
    const int g = ch / half, ci = ch - g * half;
    s_w[i] = w3d[((size_t)g * half + ci) * 27 + tap];      // weight [2][half][3][3][3]
  }
  for (int i = threadIdx.x; i < fold; i += GS_THREADS) {
    s_scale[i] = bn_scale[i];
    s_shift[i] = bn_shift[i];
  }
  __syncthreads();

  const float bias0 = b3d[0], bias1 = b3d[1];
  float* gate_f = gate + (size_t)f * hw * 2;
  for (int p = threadIdx.x; p < hw; p += GS_THREADS) {
    const int py = p / w, px = p - py * w;
    float a0 = bias0, a1 = bias1;
    for (int dt = -1; dt <= 1; ++dt) {
      const int tt = t + dt;
      if (tt < 0 || tt >= clip_len) continue;
      const T* xf = x + (size_t)(f + dt) * hw * c;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = py + dy;
        if (yy < 0 || yy >= h) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = px + dx;
          if (xx < 0 || xx >= w) continue;
          const float* wt = s_w + ((dt + 1) * 9 + (dy + 1) * 3 + (dx + 1)) * fold;
          const T* src = xf + ((size_t)yy * w + xx) * c;
          for (int ch = 0; ch < fold; ch += 4) {          // fold % 4 == 0 and half % 2 == 0
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = fmaxf(fmaf(Elem<T>::ld(src + ch + j), s_scale[ch + j], s_shift[ch + j]), 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (ch + j < half) a0 = fmaf(v[j], wt[ch + j], a0);
              else a1 = fmaf(v[j], wt[ch + j], a1);
            }
          }
        }
      }
    }
    gate_f[2 * p] = tanhf(a0);
    gate_f[2 * p + 1] = tanhf(a1);
  }
  __syncthreads();

  // spatial sums of y = gate*x and r = x - y per channel (fixed order -> deterministic)
  const int SEG = GS_THREADS / fold > 0 ? GS_THREADS / fold : 1;
  const T* xf = x + (size_t)f * hw * c;
  for (int q = threadIdx.x; q < fold * SEG; q += GS_THREADS) {
    const int ch = q % fold, seg = q / fold;
    const int g = ch / half;
    float sy = 0.f, sr = 0.f;
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

// ---- kernel 2: GSF fusion weights.  One CTA per clip. ----
__global__ void __launch_bounds__(GS_THREADS)
gsf_weight_kernel(const float* __restrict__ sums, int clip_len, int hw, int fold, const float* __restrict__ cc_w,
                  const float* __restrict__ cc_b, float* __restrict__ wgt) {
  const int b = blockIdx.x;
  const int half = fold / 2;
  const float inv = 1.f / (float)hw;
  for (int i = threadIdx.x; i < clip_len * fold; i += GS_THREADS) {
    const int ch = i % fold, t = i / fold;
    const int g = ch / half, ci = ch - g * half;
    const float* wk = cc_w + g * 18;          // [2 (y, r)][3 (channel)][3 (time)]
    float a = cc_b[g];
    for (int dc = -1; dc <= 1; ++dc) {
      const int cc = ci + dc;
      if (cc < 0 || cc >= half) continue;
      for (int dt = -1; dt <= 1; ++dt) {
        const int tt = t + dt;
        if (tt < 0 || tt >= clip_len) continue;
        // plane 0: mean of the SHIFTED y at time tt (= y at tt+1 for g=0, tt-1 for g=1, zero outside)
        const int ts = (g == 0) ? tt + 1 : tt - 1;
        float ym = 0.f;
        if (ts >= 0 && ts < clip_len) ym = sums[(((size_t)b * clip_len + ts) * fold + g * half + cc) * 2] * inv;
        const float rm = sums[(((size_t)b * clip_len + tt) * fold + g * half + cc) * 2 + 1] * inv;
        a = fmaf(wk[(dc + 1) * 3 + (dt + 1)], ym, a);
        a = fmaf(wk[9 + (dc + 1) * 3 + (dt + 1)], rm, a);
      }
    }
    wgt[((size_t)b * clip_len + t) * fold + ch] = sigmoidf_(a);
  }
}

// ---- kernel 3: blend + channel interleave.  Thread per (pixel, output channel quad). ----
template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gsf_blend_kernel(const T* __restrict__ x, int clip_len, int hw, int c, int fold, int mode,
                 const float* __restrict__ gate, const float* __restrict__ wgt, T* __restrict__ out, int ld_out,
                 long long total) {
  const long long idx = (long long)blockIdx.x * GS_THREADS + threadIdx.x;
  if (idx >= total) return;
  const int half = fold / 2, quarter = fold / 4;
  const int jo = (int)(idx % ld_out);               // output (interleaved) channel, or a pad column
  const long long fp = idx / ld_out;                // frame*hw + pixel
  if (jo >= fold) {                                 // pad columns feed zero weights in the GEMM: keep them finite
    Elem<T>::st(out + (size_t)fp * ld_out + jo, 0.f);
    return;
  }
  const int p = (int)(fp % hw);
  const long long f = fp / hw;
  const int t = (int)(f % clip_len);
  const int g = jo / half, jj = jo - g * half;
  const int ch = g * half + (jj & 1) * quarter + (jj >> 1);   // out[2i+k] = in[k*quarter + i]
  const float xv = Elem<T>::ld(x + ((size_t)f * hw + p) * c + ch);
  const float gv = gate[((size_t)f * hw + p) * 2 + g];
  const float r = xv - gv * xv;
  const int dt = (g == 0) ? 1 : -1;
  float ys = 0.f;
  if (t + dt >= 0 && t + dt < clip_len) {
    const long long fs = f + dt;
    ys = gate[((size_t)fs * hw + p) * 2 + g] * Elem<T>::ld(x + ((size_t)fs * hw + p) * c + ch);
  }
  float o;
  if (mode == TDEED_SHIFT_GSF) {
    const float wv = wgt[(size_t)f * fold + ch];
    o = ys * wv + r * (1.f - wv);
  } else {
    o = ys + r;
  }
  Elem<T>::st(out + ((size_t)f * hw + p) * ld_out + jo, o);
}

template <typename T>
static int launch_gsf(int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                      const float* bn_scale, const float* bn_shift, const float* w3d, const float* b3d,
                      const float* cc_w, const float* cc_b, float* ws, void* out, int ld_out, cudaStream_t st) {
  const int n = clips * clip_len, hw = h * w;
  float* gate = ws;
  float* sums = gate + (size_t)n * hw * 2;
  float* wgt = sums + (size_t)n * fold * 2;
  const int SEG = GS_THREADS / fold > 0 ? GS_THREADS / fold : 1;
  const size_t smem = ((size_t)27 * fold + 2 * fold + (size_t)SEG * fold * 2) * sizeof(float);
  gsf_gate_kernel<T><<<n, GS_THREADS, smem, st>>>((const T*)x, clip_len, h, w, c, fold, bn_scale, bn_shift, w3d, b3d,
                                                  gate, sums);
  int rc = check_launch("tdeed_gsf_fwd(gate)");
  if (rc) return rc;
  if (mode == TDEED_SHIFT_GSF) {
    gsf_weight_kernel<<<clips, GS_THREADS, 0, st>>>(sums, clip_len, hw, fold, cc_w, cc_b, wgt);
    rc = check_launch("tdeed_gsf_fwd(weights)");
    if (rc) return rc;
  }
  const long long total = (long long)n * hw * ld_out;
  gsf_blend_kernel<T><<<(unsigned)ceil_div_ll(total, GS_THREADS), GS_THREADS, 0, st>>>(
      (const T*)x, clip_len, hw, c, fold, mode, gate, wgt, (T*)out, ld_out, total);
  return check_launch("tdeed_gsf_fwd(blend)");
}

}  // namespace tdeed

extern "C" long long tdeed_gsf_workspace_floats(int clips, int clip_len, int h, int w, int fold) {
  const long long n = (long long)clips * clip_len;
  return n * h * w * 2 + n * fold * 2 + n * fold;
}

extern "C" int tdeed_gsf_fwd(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                             const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                             const float* cc_w, const float* cc_b, float* workspace, void* out, int ld_out,
                             void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && bn_scale && bn_shift && conv3d_w && conv3d_b && workspace && out, TDEED_ERR_SHAPE,
                "tdeed_gsf_fwd: null pointer");
  TDEED_REQUIRE(mode == TDEED_SHIFT_GSM || (cc_w && cc_b), TDEED_ERR_SHAPE, "tdeed_gsf_fwd: GSF needs channel_conv weights");
  TDEED_REQUIRE(clips > 0 && clip_len > 0 && h > 0 && w > 0 && fold > 0 && fold % 4 == 0 && fold <= c && fold <= 1024 &&
                ld_out >= fold, TDEED_ERR_SHAPE,
                "tdeed_gsf_fwd: bad shape clips=%d T=%d %dx%dx%d fold=%d ld_out=%d", clips, clip_len, h, w, c, fold, ld_out);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16)
    return launch_gsf<__nv_bfloat16>(mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b,
                                     cc_w, cc_b, workspace, out, ld_out, st);
  if (dtype == TDEED_F32)
    return launch_gsf<float>(mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                             workspace, out, ld_out, st);
  set_error("tdeed_gsf_fwd: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}
