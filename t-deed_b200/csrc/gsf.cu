// (5) Gate-Shift (GSM) / Gate-Shift-Fuse (GSF) on the first `fold` channels of an NHWC activation.
// Reference: model/shift.py:64-93, model/impl/gsm.py:89-116, model/impl/gsf.py:38-93 (eval-mode BN3d).
//
//   gate[b,g,t,p] = tanh( conv3d_{3x3x3, pad 1}( relu(bn(x)) )[group g] )
//   y = gate * x, r = x - y;  spatial sums of y and r per (frame, channel)
//   GSF: w[b,c,t] = sigmoid( conv2d_{2->1,3x3,pad1} over the (channel,time) plane of [mean(shift(y)), mean(r)] )
//   out = shift(y)*w + r*(1-w)   (GSM: shift(y) + r), channel-interleaved
//   shift: group 0 takes y from t+1 (zero at T-1), group 1 from t-1 (zero at 0); no leakage across clips.
//
// bf16 inference takes the tensor-core path further down: gsf_pack_w_kernel -> gsf_gate_tc_kernel (gate conv as an mma.sync implicit
// GEMM fused with tanh and the y / x sums) -> gsf_weight_kernel -> gsf_blend8n_kernel / gsf_blend8_kernel (16-byte pieces).  The
// fp32 exact mode and shapes the tensor-core plan rejects use the CUDA-core kernels below; the training forward uses the
// tensor-core gate kernel and the CUDA-core blend (it also copies the untouched channels).
//
// CUDA-core kernel plan (every input frame is read from HBM once per kernel, all reductions in a fixed order):
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
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace tdeed {

constexpr int GS_THREADS = 256;

// the blend arithmetic, spelled with explicit roundings so that every blend kernel gives the same bits
__device__ __forceinline__ float gs_resid(float x, float g) { return __fmaf_rn(-g, x, x); }                    // r = x - g*x
__device__ __forceinline__ float gs_fuse(float ys, float r, float w) { return __fmaf_rn(ys, w, __fmul_rn(r, __fsub_rn(1.f, w))); }
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
                  const float* __restrict__ cc_b, float* __restrict__ wgt, int RB) {
  // sums: [frame][RB row blocks][fold][2]; the row-block partials (tensor-core gate kernel) are added in block order
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
        float ym = 0.f, rm = 0.f;
        for (int r = 0; r < RB; ++r) {
          if (ts >= 0 && ts < clip_len) ym += sums[((((size_t)(f - t + ts)) * RB + r) * fold + g * half + cc) * 2];
          rm += sums[((((size_t)(f - t + tt)) * RB + r) * fold + g * half + cc) * 2 + 1];
        }
        ym *= inv;
        rm *= inv;
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
                 long long total, int copy_tail, int natural) {
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
      const int ch = natural ? jo : g * half + (jj & 1) * quarter + (jj >> 1);   // out[2i+k] = in[k*quarter + i]
      const float xv = Elem<T>::ld(xt + ch);
      const float r = gs_resid(xv, g == 0 ? g0 : g1);
      float ys = 0.f;
      if (g == 0) { if (has_next) ys = __fmul_rn(gn, Elem<T>::ld(xn + ch)); }
      else { if (has_prev) ys = __fmul_rn(gp, Elem<T>::ld(xp + ch)); }
      if (mode == TDEED_SHIFT_GSF) {
        const float wv = wgt[(size_t)f * fold + ch];
        v = gs_fuse(ys, r, wv);
      } else {
        v = __fadd_rn(ys, r);
      }
    }
    else if (copy_tail && jo < c) v = Elem<T>::ld(xt + jo);   // training: materialise the concat [gs(x[:, :fold]) | x[:, fold:]]
    o[j] = v;
  }
  store8(out + (size_t)fp * ld_out + o8 * 8, o);
}

// ---- kernels 1+2 fused on tensor cores (bf16 inference): gate = tanh(conv3d(relu(bn(x)))) + the per-channel sums of y and r ----
// The 3x3x3 gate convolution as an implicit GEMM on mma.sync m16n8k16 (bf16 operands, fp32 accumulate):
//   rows   p' : the positions of the CTA's row block INCLUDING the two halo columns (R x (w+2)), 16 per MMA
//   K         : the fold channels (padded to 16) of z = relu(bn(x)), staged pixel-major in shared memory (pitch = K*2+16 bytes:
//               ldmatrix rows of consecutive positions fall into distinct 16-byte bank groups)
//   N = 8     : (group g, horizontal tap dx) -> 6 columns; the weights of the other group's channels are zero
//   P_dx[p']  = sum over (source frame dt, vertical tap dy, channel) of W[g][ch][dt][dy][dx] * z[t+dt-1][p' + dy*(w+2)][ch]
// so the temporal and vertical taps are A-operand address offsets into the halo tile and only the horizontal tap is left as a
// shift-add of three neighbouring P columns: gate[g][y][x] = tanh(b[g] + sum_dx P_dx[y*(w+2) + x + dx][g*3+dx]).
// One CTA owns FT consecutive frames of a clip x R rows: the FT+2 staged frames are shared by its outputs (the CUDA-core
// gsf_q_kernel + gsf_gate_kernel pair ran at 187/260/397 us for the rny002 layers, L1/issue-bound; no Q round trip here).
constexpr int GT_THREADS = 256;                // 512 threads doubled the per-thread set-up work: 244M vs 164M instructions, slower
constexpr int GT_SMEM_BUDGET = 110 * 1024;     // two CTAs per SM

__device__ __forceinline__ void gt_ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void gt_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// totals of the row-block partial sums (training: the backward pass reads [frame][fold][2]), blocks added in order
__global__ void gsf_sum_partials_kernel(const float* __restrict__ part, int n_frames, int RB, int per_frame, float* __restrict__ sums) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_frames * per_frame) return;
  const int f = idx / per_frame, i = idx - f * per_frame;
  float a = 0.f;
  for (int r = 0; r < RB; ++r) a += part[((size_t)f * RB + r) * per_frame + i];
  sums[idx] = a;
}

// B fragments of the gate conv, one uint2 per (dt, dy, k-step, lane): [k = s*16 + (lane%4)*2 + {0,1} (+8)][n = lane/4]
__global__ void gsf_pack_w_kernel(const float* __restrict__ w3d, int fold, int ks, uint2* __restrict__ wB) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * ks * 32) return;
  const int lane = idx & 31, s = (idx >> 5) % ks, tap = (idx >> 5) / ks;      // tap = dt*3 + dy
  const int n = lane >> 2, kb = s * 16 + (lane & 3) * 2;
  const int half = fold >> 1;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int ch = kb + (q & 1) + (q >> 1) * 8;
    v[q] = 0.f;
    if (n < 6 && ch < fold) {
      const int g = n / 3, dx = n - g * 3;
      if ((ch >= half ? 1 : 0) == g) v[q] = w3d[(size_t)ch * 27 + tap * 3 + dx];   // [g][ci][kt][dy][dx], ch = g*half + ci
    }
  }
  wB[idx] = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

constexpr int GT_NI = 32;                      // halo positions per thread (one validity bit each)

// KS = k-steps (fold padded to 16 channels per step); KS = 0: run-time `ks`
template <int KS>
__global__ void __launch_bounds__(GT_THREADS, 2)
gsf_gate_tc_kernel(const __nv_bfloat16* __restrict__ x, int clip_len, int h, int w, int c, int fold, int ks_rt, int R, int FT, int MT,
                   int z_bytes, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, const uint2* __restrict__ wB,
                   const float* __restrict__ b3d, float* __restrict__ gate, float* __restrict__ sums_part, int RB) {
  extern __shared__ __align__(128) unsigned char gt_smem[];
  const int ks = KS ? KS : ks_rt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wp = w + 2, hpx = (R + 2) * wp, hw = h * w;
  const int foldp = ks * 16, pitch = foldp * 2 + 16, nchunk = foldp >> 3;
  uint2* s_wB = reinterpret_cast<uint2*>(gt_smem);
  unsigned char* s_z = gt_smem + 9 * ks * 256;
  float* s_P = reinterpret_cast<float*>(s_z + z_bytes);
  float2* s_gate = reinterpret_cast<float2*>(s_P + FT * MT * 128);
  const int rb = blockIdx.x, b = blockIdx.z;
  const int y0 = rb * R, t0 = blockIdx.y * FT;
  const int nfo = min(FT, clip_len - t0);
  const int rows = min(R, h - y0);
  const size_t f0 = (size_t)b * clip_len + t0;          // first output frame
  const uint32_t z_addr = (uint32_t)__cvta_generic_to_shared(s_z);
  const uint32_t frame_bytes = (uint32_t)(hpx * pitch);

  // ---- stage frames t0-1 .. t0+nfo, rows y0-1 .. y0+R, columns -1 .. w: raw x by cp.async (every 16-byte piece of the CTA is in
  //      flight at once), zeros outside the image / the clip, then z = relu(bn(x)) in place on the pieces that hold pixels.
  //      A thread visits the same <= GT_NI halo positions in every frame: their global offsets are computed once. ----
  const int k = tid % nchunk, rl = tid / nchunk, rpp = GT_THREADS / nchunk;
  const int ch0 = k * 8;
  const bool chunk_ok = ch0 + 8 <= c;
  unsigned gmask = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) gmask |= (unsigned)(ch0 + j >= (fold >> 1) ? 1 : 0) << j;
  const uint32_t zt = z_addr + (uint32_t)(rl * pitch + k * 16);
  const uint32_t step = (uint32_t)(rpp * pitch);
  const int step_y = rpp / wp, step_x = rpp - step_y * wp;
  const int nfr = nfo + 2;
  unsigned fmask = 0;                                   // staged frames that lie inside the clip
  for (int fr = 0; fr < nfr; ++fr) fmask |= (unsigned)(t0 - 1 + fr >= 0 && t0 - 1 + fr < clip_len ? 1 : 0) << fr;
  unsigned vmask = 0;                                   // this thread's halo positions (<= 32) that hold a pixel
  if (rl < rpp) {
    const __nv_bfloat16* xc = x + ((size_t)b * clip_len + t0 - 1) * hw * c + ch0;     // frame fr: xc + fr * hw * c (fmask-guarded)
    int hy = rl / wp, hx = rl - hy * wp, i = 0;
    uint32_t zp = zt;
    for (int hp = rl; hp < hpx; hp += rpp, ++i, zp += step) {
      const int py = y0 - 1 + hy, px = hx - 1;
      const bool pv = chunk_ok && py >= 0 && py < h && px >= 0 && px < w;
      vmask |= (unsigned)(pv ? 1 : 0) << i;
      const __nv_bfloat16* src = xc + (pv ? (py * w + px) * c : 0);
      uint32_t zq = zp;
      for (int fr = 0; fr < nfr; ++fr, zq += frame_bytes, src += (size_t)hw * c) {
        const bool in = pv && ((fmask >> fr) & 1u);     // src-size 0 = zero fill: one instruction, no divergent branch
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(zq), "l"(in ? (const void*)src : (const void*)x),
                     "r"(in ? 16 : 0) : "memory");
      }
      hy += step_y; hx += step_x;
      if (hx >= wp) { hx -= wp; ++hy; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  {                                                     // the packed weights ride in a second group of the same wait
    const uint32_t wb_addr = (uint32_t)__cvta_generic_to_shared(s_wB);
    for (int i = tid; i < 9 * ks * 16; i += GT_THREADS)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wb_addr + (uint32_t)i * 16u), "l"(reinterpret_cast<const uint4*>(wB) + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = ch0 + j;
    sc[j] = ch < fold ? bn_scale[ch] : 0.f;              // pad channels: z = relu(0 * x + 0) = 0
    sh[j] = ch < fold ? bn_shift[ch] : 0.f;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");     // every thread: the weight pieces of the idle tail threads too
  {
    uint32_t zp = zt;
    for (unsigned m = vmask; m; m >>= 1, zp += step) {
      if (!(m & 1u)) continue;
      uint32_t zq = zp;
      for (unsigned fm = fmask; fm; fm >>= 1, zq += frame_bytes) {
        if (!(fm & 1u)) continue;                         // this thread's own copies: visible after wait_group
        uint32_t wv[4], o[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(wv[0]), "=r"(wv[1]), "=r"(wv[2]), "=r"(wv[3]) : "r"(zq));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float lo = fmaxf(fmaf(__uint_as_float(wv[q] << 16), sc[2 * q], sh[2 * q]), 0.f);
          const float hi = fmaxf(fmaf(__uint_as_float(wv[q] & 0xffff0000u), sc[2 * q + 1], sh[2 * q + 1]), 0.f);
          o[q] = pack_bf16x2(lo, hi);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(zq), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
      }
    }
  }
  __syncthreads();

  // ---- implicit GEMM: a warp per (output frame, 16-position tile) ----
  {
    const uint32_t b_addr = (uint32_t)__cvta_generic_to_shared(s_wB) + (uint32_t)lane * 8u;
    const uint32_t dy_bytes = (uint32_t)(wp * pitch);
    for (int item = warp; item < nfo * MT; item += GT_THREADS / 32) {
      const int fo = item / MT, mt = item - fo * MT;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t a_base = z_addr + (uint32_t)fo * frame_bytes + (uint32_t)((mt * 16 + (lane & 15)) * pitch) + (uint32_t)(lane >> 4) * 16u;
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const uint32_t a_addr = a_base + (uint32_t)dt * frame_bytes + (uint32_t)dy * dy_bytes;
          const uint32_t bq = b_addr + (uint32_t)((dt * 3 + dy) * ks) * 256u;
          if (KS) {
#pragma unroll
            for (int s2 = 0; s2 < KS; ++s2) {
              uint32_t a[4], b0, b1;
              gt_ldmatrix_x4(a_addr + (uint32_t)s2 * 32u, a);
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(bq + (uint32_t)s2 * 256u));
              gt_mma16816(acc, a, b0, b1);
            }
          } else {
#pragma unroll 2
            for (int s2 = 0; s2 < ks; ++s2) {
              uint32_t a[4], b0, b1;
              gt_ldmatrix_x4(a_addr + (uint32_t)s2 * 32u, a);
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(bq + (uint32_t)s2 * 256u));
              gt_mma16816(acc, a, b0, b1);
            }
          }
        }
      }
      float* P = s_P + item * 128 + (lane >> 2) * 8 + (lane & 3) * 2;
      *reinterpret_cast<float2*>(P) = make_float2(acc[0], acc[1]);
      *reinterpret_cast<float2*>(P + 64) = make_float2(acc[2], acc[3]);
    }
  }
  __syncthreads();

  // ---- gate = tanh(bias + horizontal shift-add) ----
  {
    const float bias0 = b3d[0], bias1 = b3d[1];
    const int npx = rows * w;
    for (int i = tid; i < nfo * npx; i += GT_THREADS) {
      const int fo = i / npx, rem = i - fo * npx;
      const int y = rem / w, xx = rem - y * w;
      const float* P = s_P + fo * MT * 128 + (y * wp + xx) * 8;
      const float a0 = bias0 + P[0] + P[8 + 1] + P[16 + 2];
      const float a1 = bias1 + P[3] + P[8 + 4] + P[16 + 5];
      const float2 g = make_float2(tanhf(a0), tanhf(a1));
      *reinterpret_cast<float2*>(gate + ((f0 + fo) * hw + (size_t)y0 * w + rem) * 2) = g;
      s_gate[fo * R * w + rem] = g;
    }
  }
  __syncthreads();

  // ---- spatial sums of y = gate*x and of x per (frame, channel) -> (sum y, sum r = sum x - sum y): a thread owns (frame, 8 channels,
  //      pixel slice), the S slices of an output are added in a fixed order ----
  {
    float* s_part = reinterpret_cast<float*>(s_z);        // [nfo * S][foldp][2] (the z tile is dead)
    const int npx = rows * w;
    const int slots = nfo * nchunk;
    const int S = GT_THREADS / slots;                     // >= 1: host keeps FT * nchunk <= GT_THREADS
    const int slot = tid / S, sl = tid - slot * S;
    if (slot < slots) {
      const int fo = slot / nchunk, kk = slot - fo * nchunk;
      const int c0 = kk * 8;
      float sy[8], sx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sy[j] = sx[j] = 0.f;
      if (c0 + 8 <= c) {
        const int half = fold >> 1;
        const __nv_bfloat16* xf = x + ((f0 + fo) * hw + (size_t)y0 * w) * c + c0;
        const float2* gs = s_gate + fo * R * w;
        const bool lo_only = c0 + 8 <= half, hi_only = c0 >= half;
        for (int p = sl; p < npx; p += 2 * S) {           // two pixels per round: both loads in flight
          const int p1 = p + S;
          const uint4 r0 = *reinterpret_cast<const uint4*>(xf + (uint32_t)(p * c));
          uint4 r1 = make_uint4(0u, 0u, 0u, 0u);
          float2 g1 = make_float2(0.f, 0.f);
          if (p1 < npx) { r1 = *reinterpret_cast<const uint4*>(xf + (uint32_t)(p1 * c)); g1 = gs[p1]; }
          const float2 g0 = gs[p];
          const uint32_t w0[4] = {r0.x, r0.y, r0.z, r0.w}, w1[4] = {r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v0 = (j & 1) ? __uint_as_float(w0[j >> 1] & 0xffff0000u) : __uint_as_float(w0[j >> 1] << 16);
            const float v1 = (j & 1) ? __uint_as_float(w1[j >> 1] & 0xffff0000u) : __uint_as_float(w1[j >> 1] << 16);
            const bool hi = hi_only || (!lo_only && c0 + j >= half);
            sy[j] = fmaf(hi ? g0.y : g0.x, v0, sy[j]);
            sx[j] += v0;
            sy[j] = fmaf(hi ? g1.y : g1.x, v1, sy[j]);      // zeros when p1 is past the block
            sx[j] += v1;
          }
        }
      }
      float2* dst = reinterpret_cast<float2*>(s_part) + (fo * S + sl) * foldp + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float2(sy[j], sx[j] - sy[j]);
    }
    __syncthreads();
    // outputs (frame, channel, {y, r}); G threads (a power of two: contiguous lanes of one warp) add strided slices, then a shuffle tree
    const int nout = nfo * fold * 2;
    int G = 1;
    while (G < 32 && G * 2 * nout <= GT_THREADS && G * 2 <= S) G *= 2;
    for (int o0 = 0; o0 < nout * G; o0 += GT_THREADS) {
      const int og = o0 + tid;
      const int o = og / G, sub = og - o * G;
      const int fo = o / (fold * 2), i = o - fo * (fold * 2);
      float a = 0.f;
      if (o < nout) {
        const float* sp = s_part + fo * S * foldp * 2 + i;
        const int stride = foldp * 2;
#pragma unroll 4
        for (int q = sub; q < S; q += G) a += sp[q * stride];
      }
      for (int sh2 = 1; sh2 < G; sh2 <<= 1) a += __shfl_xor_sync(0xffffffffu, a, sh2);
      if (o < nout && sub == 0) sums_part[(((f0 + fo) * RB + rb) * fold) * 2 + i] = a;
    }
  }
}

// ---- kernel 4b (bf16 inference): blend + channel interleave on 16-byte pieces.  One CTA per `subs` tiles of GSB_ROWS (frame, pixel) rows ----
// A thread owns 8 consecutive INPUT channels (the same 8 for every row it visits, so the interleaved output positions and the
// group tests are computed once): per row one 16-byte load of x[t] and of the shifted neighbour frame(s), eight blends, eight
// 2-byte stores into the CTA's [rows][ld_out] shared-memory tile; the tile (a contiguous piece of `out`) leaves with coalesced
// 16-byte stores.  (gsf_blend_kernel walks the OUTPUT channels with 2-byte global loads and index arithmetic per channel:
// ~500 instructions per 8 channels, instruction-bound at ~10x the time the bytes need.)
constexpr int GSB_ROWS = 128;
__global__ void __launch_bounds__(GS_THREADS)
gsf_blend8_kernel(const __nv_bfloat16* __restrict__ x, int clip_len, int hw, int c, int fold, int mode, unsigned rows_total, int subs,
                  const float* __restrict__ gate, const float* __restrict__ wgt, __nv_bfloat16* __restrict__ out, int ld_out) {
  extern __shared__ __align__(16) unsigned char gsb_smem[];
  __nv_bfloat16* s_out = reinterpret_cast<__nv_bfloat16*>(gsb_smem);                 // [GSB_ROWS][ld_out]
  const int half = fold >> 1, quarter = fold >> 2;
  const int nchunk = ld_out >> 3;
  const int rpp = GS_THREADS / nchunk;                    // rows per pass
  const int k = threadIdx.x % nchunk, rl = threadIdx.x / nchunk;
  const int ch0 = k * 8;
  // per-thread constants: output position of each of the 8 channels, its group
  int jo[8];
  unsigned gmask = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = ch0 + j;
    const int g = ch >= half ? 1 : 0;
    const int ci = ch - g * half;
    const int k2 = ci >= quarter ? 1 : 0;
    jo[j] = ch < fold ? g * half + 2 * (ci - k2 * quarter) + k2 : ch;    // out[2i+k] = in[k*quarter + i]; pad columns keep their place
    gmask |= (unsigned)g << j;
  }
  const bool lo_grp = ch0 < half, hi_grp = ch0 + 7 >= half;
  const int nvalid = min(8, fold - ch0);                  // channels >= fold are pad columns (written as zeros)
  const unsigned fstride = (unsigned)hw * (unsigned)c;    // host: hw * c and frames * hw < 2^31
  for (int sub = 0; sub < subs; ++sub) {
    const unsigned r0 = ((unsigned)blockIdx.x * (unsigned)subs + sub) * GSB_ROWS;
    if (r0 >= rows_total) break;
    const int nrows = (int)min((unsigned)GSB_ROWS, rows_total - r0);
    if (sub) __syncthreads();                             // the previous tile has left shared memory
    if (rl < rpp) {
      const unsigned r = r0 + rl;
      int f = (int)(r / (unsigned)hw);
      int p = (int)(r - (unsigned)f * (unsigned)hw);
      int t = f % clip_len;
      for (int row = rl; row < nrows; row += rpp) {
        const unsigned fp = (unsigned)f * (unsigned)hw + (unsigned)p;
        const __nv_bfloat16* xt = x + (size_t)fp * c + ch0;
        const float* gt = gate + (size_t)fp * 2;
        const bool has_next = t + 1 < clip_len, has_prev = t > 0;
        float xv[8], xn[8], xp[8], wv[8];
        load8(xt, xv);
        const float2 g01 = *reinterpret_cast<const float2*>(gt);
        float gn = 0.f, gp = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) xn[j] = xp[j] = wv[j] = 0.f;
        if (lo_grp && has_next) { load8(xt + fstride, xn); gn = gt[2 * hw]; }
        if (hi_grp && has_prev) { load8(xt - fstride, xp); gp = gt[1 - 2 * hw]; }
        if (mode == TDEED_SHIFT_GSF) {
          const float* wp = wgt + (size_t)f * fold + ch0;   // 16-byte aligned: fold % 4 == 0
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
          if (nvalid > 4) {
            const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
            wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
          }
        }
        __nv_bfloat16* so = s_out + row * ld_out;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool g1 = (gmask >> j) & 1u;
          const float r_ = gs_resid(xv[j], g1 ? g01.y : g01.x);
          const float ys = g1 ? __fmul_rn(gp, xp[j]) : __fmul_rn(gn, xn[j]);
          float v = mode == TDEED_SHIFT_GSF ? gs_fuse(ys, r_, wv[j]) : __fadd_rn(ys, r_);
          if (j >= nvalid) v = 0.f;                         // pad columns: the GEMM multiplies them by zero weights
          so[jo[j]] = __float2bfloat16_rn(v);
        }
        p += rpp;
        while (p >= hw) { p -= hw; ++f; if (++t == clip_len) t = 0; }
      }
    }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(s_out);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)r0 * ld_out);
    for (int i = threadIdx.x; i < nrows * nchunk; i += GS_THREADS) dst[i] = src[i];
  }
}

// ---- kernel 4c (bf16 inference, natural channel order): out[:, ch] = blend of input channel ch.  The reference's channel
// interleave (out[2i+k] = in[k*quarter + i]) is a fixed permutation in front of a 1x1 convolution: the caller folds it into the
// columns of that convolution's weight, and a thread's 8 input channels leave as one 16-byte store (no shared-memory tile,
// no per-channel output positions). ----
__global__ void __launch_bounds__(GS_THREADS)
gsf_blend8n_kernel(const __nv_bfloat16* __restrict__ x, int clip_len, int hw, int c, int fold, int mode, unsigned rows_total, int iters,
                   const float* __restrict__ gate, const float* __restrict__ wgt, __nv_bfloat16* __restrict__ out, int ld_out) {
  const int nchunk = ld_out >> 3, rpp = GS_THREADS / nchunk;
  const int k = threadIdx.x % nchunk, rl = threadIdx.x / nchunk;
  if (rl >= rpp) return;
  const int ch0 = k * 8, half = fold >> 1;
  const bool lo_grp = ch0 < half, hi_grp = ch0 + 7 >= half;
  const int nvalid = min(8, fold - ch0);                  // channels >= fold are pad columns (written as zeros)
  unsigned r = blockIdx.x * (unsigned)(rpp * iters) + rl; // row = frame * hw + pixel (host: frames * hw, hw * c < 2^31)
  if (r >= rows_total) return;
  int f = (int)(r / (unsigned)hw);
  int p = (int)(r - (unsigned)f * (unsigned)hw);
  int t = f % clip_len;
  const unsigned fstride = (unsigned)hw * (unsigned)c;
  for (int it = 0; it < iters && r < rows_total; ++it, r += rpp) {
    const __nv_bfloat16* xt = x + (size_t)r * c + ch0;
    const float* gt = gate + (size_t)r * 2;
    const bool has_next = t + 1 < clip_len, has_prev = t > 0;
    float xv[8], xn[8], xp[8], wv[8];
    load8(xt, xv);
    const float2 g01 = *reinterpret_cast<const float2*>(gt);
    float gn = 0.f, gp = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { xn[j] = xp[j] = 0.f; wv[j] = 1.f; }
    if (lo_grp && has_next) { load8(xt + fstride, xn); gn = gt[2 * hw]; }
    if (hi_grp && has_prev) { load8(xt - fstride, xp); gp = gt[1 - 2 * hw]; }
    if (mode == TDEED_SHIFT_GSF) {
      const float* wp = wgt + (size_t)f * fold + ch0;     // 16-byte aligned: fold % 4 == 0
      const float4 w0 = *reinterpret_cast<const float4*>(wp);
      wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
      if (nvalid > 4) {
        const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
        wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
      }
    }
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool g1 = !lo_grp || (hi_grp && ch0 + j >= half);
      const float r_ = gs_resid(xv[j], g1 ? g01.y : g01.x);
      const float ys = g1 ? __fmul_rn(gp, xp[j]) : __fmul_rn(gn, xn[j]);
      v[j] = mode == TDEED_SHIFT_GSF ? gs_fuse(ys, r_, wv[j]) : __fadd_rn(ys, r_);
      if (j >= nvalid) v[j] = 0.f;                        // pad columns: the GEMM multiplies them by zero weights
    }
    store8(out + (size_t)r * ld_out + ch0, v);
    p += rpp;
    while (p >= hw) { p -= hw; ++f; if (++t == clip_len) t = 0; }
  }
}

template <typename T>
static int launch_gsf(int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                      const float* bn_scale, const float* bn_shift, const float* w3d, const float* b3d,
                      const float* cc_w, const float* cc_b, float* ws, void* out, int ld_out, int copy_tail, int natural, cudaStream_t st) {
  const int n = clips * clip_len, hw = h * w;
  TDEED_REQUIRE(n <= 65535, TDEED_ERR_UNSUPPORTED, "tdeed_gsf_fwd: %d frames per call (the Q-map kernel puts frames on grid.y: at most 65535; split the batch)", n);
  float* gate = ws;
  float* sums = gate + (size_t)n * hw * 2;
  float* wgt = sums + (size_t)n * fold * 2;
  float* Q = wgt + (size_t)n * fold;
  Q += (4 - ((Q - ws) & 3)) & 3;                       // 16-byte align (float2 stores need 8)

  int RB = 1;                 // row blocks whose partial sums the weight kernel adds (tensor-core gate kernel only)
  const float* sums_in = sums;
  bool gate_done = false;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    // tensor-core gate kernel (inference): plan the row block R and the frames per CTA FT inside the shared-memory budget
    const int ks = (fold + 15) / 16, foldp = ks * 16, pitch = foldp * 2 + 16, wp = w + 2;
    const int rpp = GT_THREADS / (foldp / 8);
    int R = 0, FT = 0, MT = 0;
    size_t z_bytes = 0, smem_tc = 0;
    static int gt_budget = -1, gt_ftmax = 4;
    if (gt_budget < 0) {
      const char* e1 = tdeed::dev_env("TDEED_GT_BUDGET_KB");
      const char* e2 = tdeed::dev_env("TDEED_GT_FTMAX");
      gt_budget = e1 ? atoi(e1) * 1024 : GT_SMEM_BUDGET;
      gt_ftmax = e2 ? atoi(e2) : 4;
    }
    auto plan = [&](int rb, int ft) {
      const int r = ceil_div(h, rb), mt = ceil_div(r * wp, 16);
      size_t zb = ((size_t)(ft + 2) * (r + 2) * wp + 16) * pitch;
      const size_t part = (size_t)GT_THREADS * 64;             // [FT * S slices][foldp][2] floats, FT * S * (foldp / 8) <= GT_THREADS
      if (ft * (foldp / 8) > GT_THREADS) return false;
      if (zb < part) zb = part;
      zb = (zb + 15) / 16 * 16;
      const size_t total = (size_t)9 * ks * 256 + zb + (size_t)ft * mt * 512 + (size_t)ft * r * w * 8;
      if (total > (size_t)gt_budget || ceil_div((r + 2) * wp, rpp) > GT_NI) return false;
      RB = ceil_div(h, r); R = r; FT = ft; MT = mt; z_bytes = zb; smem_tc = total;
      return true;
    };
    bool ok = false;
    if (foldp <= 8 * GT_THREADS && clip_len >= 1 && (long long)hw * c < (1ll << 31)) {
      for (int minft = 2; minft >= 1 && !ok; --minft)
        for (int rb = 1; rb <= h && !ok; ++rb)
          for (int ft = (clip_len < gt_ftmax ? clip_len : gt_ftmax); ft >= minft && !ok; --ft) ok = plan(rb, ft);
    }
    // the partial sums of RB > 1 row blocks live in the (otherwise unused) Q region, the packed weights behind it.  The choice
    // of the path must not depend on the number of clips: a batch and its clips one by one give bit-identical results.
    if (ok && (RB == 1 || (size_t)RB * fold * 2 <= (size_t)hw * 6) && ceil_div(clip_len, FT) <= 65535 && clips <= 65535) {
      uint2* wB = reinterpret_cast<uint2*>(Q + ((size_t)n * hw * 6 + 3) / 4 * 4);
      if (tdeed::dev_env("TDEED_GT_TRACE")) fprintf(stderr, "gate_tc plan: %dx%d fold %d -> RB %d R %d FT %d MT %d smem %zu\n", h, w, fold, RB, R, FT, MT, smem_tc);
      float* sums_part = RB > 1 ? Q : sums;
      gsf_pack_w_kernel<<<ceil_div(9 * ks * 32, 128), 128, 0, st>>>(w3d, fold, ks, wB);
      int rc0 = check_launch("tdeed_gsf_fwd(pack)");
      if (rc0) return rc0;
      void (*kern)(const __nv_bfloat16*, int, int, int, int, int, int, int, int, int, int, const float*, const float*, const uint2*,
                   const float*, float*, float*, int) = gsf_gate_tc_kernel<0>;
      int kidx = 0;
      switch (ks) {                                      // the RegNetY-200MF / 800MF fold widths get unrolled k-loops
        case 1: kern = gsf_gate_tc_kernel<1>; kidx = 1; break;
        case 2: kern = gsf_gate_tc_kernel<2>; kidx = 2; break;
        case 3: kern = gsf_gate_tc_kernel<3>; kidx = 3; break;
        case 5: kern = gsf_gate_tc_kernel<5>; kidx = 4; break;
        case 6: kern = gsf_gate_tc_kernel<6>; kidx = 5; break;
        case 12: kern = gsf_gate_tc_kernel<12>; kidx = 6; break;
        default: break;
      }
      static bool tc_set[7] = {false, false, false, false, false, false, false};
      if (!tc_set[kidx]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_gsf_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        tc_set[kidx] = true;
      }
      kern<<<dim3((unsigned)RB, (unsigned)ceil_div(clip_len, FT), (unsigned)clips), GT_THREADS, smem_tc, st>>>(
          (const __nv_bfloat16*)x, clip_len, h, w, c, fold, ks, R, FT, MT, (int)z_bytes, bn_scale, bn_shift, wB, b3d, gate, sums_part, RB);
      rc0 = check_launch("tdeed_gsf_fwd(gate_tc)");
      if (rc0) return rc0;
      if (copy_tail && RB > 1) {       // training: tdeed_gsf_bwd reads the per-frame totals
        gsf_sum_partials_kernel<<<ceil_div(n * fold * 2, GS_THREADS), GS_THREADS, 0, st>>>(sums_part, n, RB, fold * 2, sums);
        rc0 = check_launch("tdeed_gsf_fwd(sum partials)");
        if (rc0) return rc0;
      }
      sums_in = sums_part;
      gate_done = true;
    } else {
      RB = 1;
    }
  }
  int rc = 0;
  if (!gate_done) {
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
  rc = check_launch("tdeed_gsf_fwd(q)");
  if (rc) return rc;

  const int SEG = GS_THREADS / fold > 0 ? GS_THREADS / fold : 1;
  gsf_gate_kernel<T><<<n, GS_THREADS, (size_t)SEG * fold * 2 * sizeof(float), st>>>((const T*)x, clip_len, hw, c, fold, b3d, Q, gate, sums);
  rc = check_launch("tdeed_gsf_fwd(gate)");
  if (rc) return rc;
  }
  if (mode == TDEED_SHIFT_GSF) {
    gsf_weight_kernel<<<n, fold < GS_THREADS ? ((fold + 31) / 32 * 32) : GS_THREADS, 0, st>>>(sums_in, clip_len, hw, fold, cc_w, cc_b, wgt, RB);
    rc = check_launch("tdeed_gsf_fwd(weights)");
    if (rc) return rc;
  }
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (natural && !copy_tail && ld_out == (fold + 7) / 8 * 8 && ld_out <= 8 * GS_THREADS && (long long)n * hw < (1ll << 31) &&
        (long long)hw * c < (1ll << 31)) {
      const unsigned rows_total = (unsigned)((long long)n * hw);
      const int rpp = GS_THREADS / (ld_out / 8), iters = 4;
      gsf_blend8n_kernel<<<(unsigned)ceil_div_ll((long long)rows_total, (long long)rpp * iters), GS_THREADS, 0, st>>>(
          (const __nv_bfloat16*)x, clip_len, hw, c, fold, mode, rows_total, iters, gate, wgt, (__nv_bfloat16*)out, ld_out);
      return check_launch("tdeed_gsf_fwd(blend8n)");
    }
    if (!natural && !copy_tail && ld_out == (fold + 7) / 8 * 8 && (size_t)GSB_ROWS * ld_out * 2 <= 48 * 1024 && (long long)n * hw < (1ll << 31) && (long long)hw * c < (1ll << 31)) {
      // tiles per CTA: the per-thread set-up is amortised over several tiles as long as ~12 CTAs per SM remain
      const unsigned rows_total = (unsigned)((long long)n * hw);
      const long long tiles = ceil_div_ll((long long)rows_total, GSB_ROWS);
      const int subs = (int)(tiles / 2000 < 1 ? 1 : (tiles / 2000 > 8 ? 8 : tiles / 2000));
      gsf_blend8_kernel<<<(unsigned)ceil_div_ll(tiles, subs), GS_THREADS, (size_t)GSB_ROWS * ld_out * 2, st>>>(
          (const __nv_bfloat16*)x, clip_len, hw, c, fold, mode, rows_total, subs, gate, wgt, (__nv_bfloat16*)out, ld_out);
      return check_launch("tdeed_gsf_fwd(blend8)");
    }
  }
  const long long total = (long long)n * hw * (ld_out / 8);
  gsf_blend_kernel<T><<<dim3((unsigned)n, (unsigned)ceil_div(hw * (ld_out / 8), GS_THREADS)), GS_THREADS, 0, st>>>(
      (const T*)x, clip_len, hw, c, fold, mode, gate, wgt, (T*)out, ld_out, total, copy_tail, natural);
  return check_launch("tdeed_gsf_fwd(blend)");
}

}  // namespace tdeed

extern "C" long long tdeed_gsf_workspace_floats(int clips, int clip_len, int h, int w, int fold) {
  const long long n = (long long)clips * clip_len;
  // gate | sums | fusion weights | Q maps (or row-block partial sums) | packed bf16 conv3d weights of the tensor-core gate kernel
  return n * h * w * 2 + n * fold * 2 + n * fold + 4 + n * h * w * 6 + 4 + 9ll * ((fold + 15) / 16) * 64;
}

static int gsf_dispatch(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                        const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                        const float* cc_w, const float* cc_b, float* workspace, void* out, int ld_out, int copy_tail,
                        int natural, void* stream) {
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
                                     cc_w, cc_b, workspace, out, ld_out, copy_tail, natural, st);
  if (dtype == TDEED_F32)
    return launch_gsf<float>(mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                             workspace, out, ld_out, copy_tail, natural, st);
  set_error("tdeed_gsf_fwd: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}

extern "C" int tdeed_gsf_fwd(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                             const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                             const float* cc_w, const float* cc_b, float* workspace, void* out, int ld_out,
                             void* stream) {
  return gsf_dispatch(dtype, mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                      workspace, out, ld_out, 0, 0, stream);
}

// same, but out[:, ch] holds input channel ch (no interleave): the caller permutes the columns of the following 1x1 convolution
// with tdeed_gsf_interleaved_position (out_natural[:, ch] == out_interleaved[:, position(ch)]).
extern "C" int tdeed_gsf_fwd_natural(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                                     const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                                     const float* cc_w, const float* cc_b, float* workspace, void* out, int ld_out,
                                     void* stream) {
  return gsf_dispatch(dtype, mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                      workspace, out, ld_out, 0, 1, stream);
}

extern "C" int tdeed_gsf_interleaved_position(int fold, int ch) {
  const int half = fold / 2, quarter = fold / 4;
  if (ch < 0 || ch >= fold || fold % 4) return -1;
  const int g = ch >= half ? 1 : 0, ci = ch - g * half, k2 = ci >= quarter ? 1 : 0;
  return g * half + 2 * (ci - k2 * quarter) + k2;         // model/impl/gsf.py:84-92: out[2i+k] = in[k*quarter + i]
}

// training: writes the full concat y = [gs(x[:, :fold]) | x[:, fold:]] as [frames*h*w, c] (model/shift.py:89-93); the
// workspace (gate, per-frame sums, fusion weights) is what tdeed_gsf_bwd reads back.
extern "C" int tdeed_gsf_cat_fwd(int dtype, int mode, const void* x, int clips, int clip_len, int h, int w, int c, int fold,
                                 const float* bn_scale, const float* bn_shift, const float* conv3d_w, const float* conv3d_b,
                                 const float* cc_w, const float* cc_b, float* workspace, void* out, void* stream) {
  return gsf_dispatch(dtype, mode, x, clips, clip_len, h, w, c, fold, bn_scale, bn_shift, conv3d_w, conv3d_b, cc_w, cc_b,
                      workspace, out, c, 1, 0, stream);
}
