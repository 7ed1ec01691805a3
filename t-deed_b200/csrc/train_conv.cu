// Backward of the spatial convolutions of the backbone (training step; the reference gets these from autograd/cuDNN):
//   tdeed_conv3x3g_bwd_data    dx = conv_transpose(dy, W)       grouped 3x3, stride 1 | 2, pad 1
//   tdeed_conv3x3g_bwd_weight  dW[co][ci][ky][kx] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, s*oy+ky-1, s*ox+kx-1, g*gw+ci]
//   tdeed_stem_bwd_weight      dW[co][ci*9+ky*3+kx] of the 3->32 stride-2 stem conv, recomputing the normalised patch
//                              from the frames (crop / flip / normalise exactly as the forward stem kernel)
// CUDA-core fp32 accumulation; weight gradients are reduced over per-CTA partials in a fixed order (deterministic).
#include <cstdlib>
#include "train_reduce.cuh"

namespace tdeed {

// ------------------------------------------------------------------------------------------------------------------
// data gradient.  Work item = (input row, strip of 4 input pixels, input-channel octet); the octet's 8 x gw x 9 weights
// sit in shared memory as [ul][ky][kx][co][ci8].
// ------------------------------------------------------------------------------------------------------------------
constexpr int CB_THREADS = 256, CB_PX = 4, CB_MAX_UNITS = 16;

template <typename T, int STRIDE>
__global__ void __launch_bounds__(CB_THREADS)
conv3x3g_bwd_data_kernel(const T* __restrict__ dy, int h, int w, int c, int gw, const float* __restrict__ weight,
                         T* __restrict__ dx, int oh, int ow, int units_per_cta) {
  extern __shared__ __align__(16) float s_w[];
  const int n_units = c / 8;
  const int u0 = blockIdx.y * units_per_cta;
  const int ucnt = min(units_per_cta, n_units - u0);
  const int per_unit = 9 * gw * 8;
  const int pitch = per_unit + 4;
  const int f = blockIdx.z;
  for (int i = threadIdx.x; i < ucnt * per_unit; i += CB_THREADS) {
    const int ul = i / per_unit, r = i - ul * per_unit;
    const int ci8 = r & 7, co = (r >> 3) % gw, tap = r / (8 * gw);
    const int u = u0 + ul;
    const int g = u * 8 / gw, ci = (u * 8) % gw + ci8;          // input channel inside its group
    s_w[ul * pitch + r] = weight[((size_t)(g * gw + co) * gw + ci) * 9 + tap];
  }
  __syncthreads();
  const int strips = (w + CB_PX - 1) / CB_PX;
  const int item = blockIdx.x * CB_THREADS + threadIdx.x;
  if (item >= h * strips * ucnt) return;
  const int ul = item % ucnt;
  const int strip = (item / ucnt) % strips;
  const int iy = item / (ucnt * strips);
  const int u = u0 + ul;
  const int co0 = (u * 8 / gw) * gw;
  const int ix0 = strip * CB_PX;
  float acc[CB_PX][8];
#pragma unroll
  for (int p = 0; p < CB_PX; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
  const T* fdy = dy + (size_t)f * oh * ow * c;
  const float* wu = s_w + ul * pitch;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ty = iy + 1 - ky;
    if (ty < 0 || (STRIDE == 2 && (ty & 1))) continue;
    const int oy = ty / STRIDE;
    if (oy >= oh) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int p = 0; p < CB_PX; ++p) {
        const int tx = ix0 + p + 1 - kx;
        if (tx < 0 || (STRIDE == 2 && (tx & 1)) || ix0 + p >= w) continue;
        const int ox = tx / STRIDE;
        if (ox >= ow) continue;
        const T* src = fdy + ((size_t)oy * ow + ox) * c + co0;
        for (int cq = 0; cq < gw; cq += 8) {
          float v[8];
          load8(src + cq, v);
#pragma unroll
          for (int co = 0; co < 8; ++co) {
            const float4* wp = reinterpret_cast<const float4*>(wu + ((ky * 3 + kx) * gw + cq + co) * 8);
            const float4 wa = wp[0], wb = wp[1];
            acc[p][0] = fmaf(v[co], wa.x, acc[p][0]);
            acc[p][1] = fmaf(v[co], wa.y, acc[p][1]);
            acc[p][2] = fmaf(v[co], wa.z, acc[p][2]);
            acc[p][3] = fmaf(v[co], wa.w, acc[p][3]);
            acc[p][4] = fmaf(v[co], wb.x, acc[p][4]);
            acc[p][5] = fmaf(v[co], wb.y, acc[p][5]);
            acc[p][6] = fmaf(v[co], wb.z, acc[p][6]);
            acc[p][7] = fmaf(v[co], wb.w, acc[p][7]);
          }
        }
      }
    }
  }
  T* fdx = dx + (size_t)f * h * w * c;
#pragma unroll
  for (int p = 0; p < CB_PX; ++p) {
    if (ix0 + p >= w) break;
    store8(fdx + ((size_t)iy * w + ix0 + p) * c + u * 8, acc[p]);
  }
}

template <typename T, int STRIDE>
static int launch_bwd_data(const void* dy, int n, int h, int w, int c, int gw, const float* weight, void* dx, cudaStream_t st) {
  const int oh = (h + STRIDE - 1) / STRIDE, ow = (w + STRIDE - 1) / STRIDE;
  const int n_units = c / 8;
  const int upc = n_units < CB_MAX_UNITS ? n_units : CB_MAX_UNITS;
  const int strips = (w + CB_PX - 1) / CB_PX;
  const size_t smem = (size_t)upc * (9 * gw * 8 + 4) * sizeof(float);
  auto kern = conv3x3g_bwd_data_kernel<T, STRIDE>;
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "conv3x3g_bwd_data: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid(ceil_div(h * strips * upc, CB_THREADS), ceil_div(n_units, upc), n);
  kern<<<grid, CB_THREADS, smem, st>>>((const T*)dy, h, w, c, gw, weight, (T*)dx, oh, ow, upc);
  return check_launch("tdeed_conv3x3g_bwd_data");
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient.  CTA = (channel block of 256/(gw*gw) groups, chunk of output-row segments); thread = one (co, ci)
// pair with its 9 taps in registers.  Per segment of <= 32 output pixels of one row the dy values and the three input
// rows are staged in shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int CW_THREADS = 256, CW_SEG = 32;

template <typename T, int STRIDE>
__global__ void __launch_bounds__(CW_THREADS)
conv3x3g_bwd_weight_kernel(const T* __restrict__ x, const T* __restrict__ dy, int n, int h, int w, int c, int gw, int oh, int ow,
                           int segs_per_row, long long total_segs, long long segs_per_cta, float* __restrict__ part) {
  constexpr int XW = (CW_SEG - 1) * STRIDE + 3;
  const int gpc = CW_THREADS / (gw * gw);          // groups per CTA
  const int cb = gpc * gw;                         // channels per CTA
  const int c0 = blockIdx.y * cb;
  const int cvalid = min(cb, c - c0);
  __shared__ float s_dy[CW_SEG][32 + 1];
  __shared__ float s_x[3][XW][32 + 1];
  const int ci = threadIdx.x % gw, co = (threadIdx.x / gw) % gw, grp = threadIdx.x / (gw * gw);
  const int lco = grp * gw + co, lci = grp * gw + ci;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const long long s_begin = (long long)blockIdx.x * segs_per_cta;
  const long long s_end = min(total_segs, s_begin + segs_per_cta);
  for (long long sidx = s_begin; sidx < s_end; ++sidx) {
    const int seg = (int)(sidx % segs_per_row);
    const long long row = sidx / segs_per_row;
    const int oy = (int)(row % oh);
    const long long f = row / oh;
    const int ox0 = seg * CW_SEG;
    const int npx = min(CW_SEG, ow - ox0);
    __syncthreads();
    for (int i = threadIdx.x; i < CW_SEG * cb; i += CW_THREADS) {
      const int ch = i % cb, p = i / cb;
      s_dy[p][ch] = (p < npx && ch < cvalid) ? Elem<T>::ld(dy + ((f * oh + oy) * ow + ox0 + p) * c + c0 + ch) : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * XW * cb; i += CW_THREADS) {
      const int ch = i % cb, px = (i / cb) % XW, ky = i / (cb * XW);
      const int iy = oy * STRIDE + ky - 1, ix = ox0 * STRIDE + px - 1;
      float v = 0.f;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w && ch < cvalid) v = Elem<T>::ld(x + ((f * h + iy) * w + ix) * c + c0 + ch);
      s_x[ky][px][ch] = v;
    }
    __syncthreads();
    if (lco < cvalid) {
      for (int p = 0; p < npx; ++p) {
        const float d = s_dy[p][lco];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc[ky * 3 + kx] = fmaf(d, s_x[ky][p * STRIDE + kx][lci], acc[ky * 3 + kx]);
      }
    }
  }
  if (lco < cvalid) {
    float* o = part + ((size_t)blockIdx.x * c + c0 + lco) * gw * 9 + ci * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) o[t] = acc[t];
  }
}

static int cw_parts(long long total_segs, int c, int gw) {
  const int cblocks = ceil_div(c, (CW_THREADS / (gw * gw)) * gw);
  long long parts = ceil_div_ll(4 * kNumSMs, cblocks);
  if (parts > total_segs) parts = total_segs;
  if (parts < 1) parts = 1;
  return (int)parts;
}

template <typename T, int STRIDE>
static int launch_bwd_weight(const void* x, const void* dy, int n, int h, int w, int c, int gw, float* dw, float* ws, cudaStream_t st) {
  const int oh = (h + STRIDE - 1) / STRIDE, ow = (w + STRIDE - 1) / STRIDE;
  const int spr = ceil_div(ow, CW_SEG);
  const long long total = (long long)n * oh * spr;
  int parts = cw_parts(total, c, gw);
  const long long spc = ceil_div_ll(total, parts);
  parts = (int)ceil_div_ll(total, spc);
  const int cb = (CW_THREADS / (gw * gw)) * gw;
  dim3 grid(parts, ceil_div(c, cb));
  conv3x3g_bwd_weight_kernel<T, STRIDE><<<grid, CW_THREADS, 0, st>>>((const T*)x, (const T*)dy, n, h, w, c, gw, oh, ow, spr, total, spc, ws);
  int rc = check_launch("tdeed_conv3x3g_bwd_weight(partial)");
  if (rc) return rc;
  const long long count = (long long)c * gw * 9;
  launch_partial_sum(ws, parts, count, dw, st);
  return check_launch("tdeed_conv3x3g_bwd_weight(final)");
}

// ------------------------------------------------------------------------------------------------------------------
// stem weight gradient.  Persistent CTAs loop over 8 x 32 output tiles; thread = (output channel, 4 of the 27 taps).
// ------------------------------------------------------------------------------------------------------------------
constexpr int SW_TW = 32, SW_TH = 8, SW_THREADS = 256;
constexpr int SW_PW = 2 * SW_TW + 1, SW_PH = 2 * SW_TH + 1;

template <typename TIn, typename TD>
__global__ void __launch_bounds__(SW_THREADS)
stem_bwd_weight_kernel(const TIn* __restrict__ frames, int unit_input, int in_h, int in_w, int crop_y, int crop_x, int h, int w,
                       int flip, const TD* __restrict__ dy, int oh, int ow, int tiles_x, int tiles_y, long long num_tiles,
                       float* __restrict__ part) {
  __shared__ float s_in[3][SW_PH][SW_PW + 1];
  __shared__ float s_dy[SW_TW * SW_TH][32 + 1];
  const int co = threadIdx.x % 32, kq = threadIdx.x / 32;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int txi = (int)(tile % tiles_x), tyi = (int)((tile / tiles_x) % tiles_y);
    const long long f = tile / ((long long)tiles_x * tiles_y);
    const int oy0 = tyi * SW_TH, ox0 = txi * SW_TW;
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    const TIn* fbase = frames + (size_t)f * 3 * in_h * in_w;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * SW_PH * SW_PW; i += SW_THREADS) {
      const int px = i % SW_PW, py = (i / SW_PW) % SW_PH, ci = i / (SW_PW * SW_PH);
      const int y = iy0 + py, x = ix0 + px;
      float v = 0.f;
      if (y >= 0 && y < h && x >= 0 && x < w) {
        const int sx = flip ? (w - 1 - x) : x;
        const float raw = (float)fbase[((size_t)ci * in_h + (crop_y + y)) * in_w + (crop_x + sx)];
        v = ((unit_input ? raw : raw / 255.f) - mean[ci]) / stdv[ci];
      }
      s_in[ci][py][px] = v;
    }
    for (int i = threadIdx.x; i < SW_TW * SW_TH * 32; i += SW_THREADS) {
      const int ch = i % 32, pix = i / 32;
      const int oy = oy0 + pix / SW_TW, ox = ox0 + pix % SW_TW;
      s_dy[pix][ch] = (oy < oh && ox < ow) ? Elem<TD>::ld(dy + ((f * oh + oy) * ow + ox) * 32 + ch) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = kq + 8 * q;
      if (k >= 27) break;
      const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
      float a = acc[q];
      for (int pix = 0; pix < SW_TW * SW_TH; ++pix) {
        const int py = pix / SW_TW, px = pix % SW_TW;
        a = fmaf(s_dy[pix][co], s_in[ci][2 * py + ky][2 * px + kx], a);
      }
      acc[q] = a;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int k = kq + 8 * q;
    if (k < 27) part[(size_t)blockIdx.x * 864 + co * 27 + k] = acc[q];
  }
}

// im2col of the normalised stem input: patches[p][k] (k = ci*9 + ky*3 + kx, 27 padded to 32 with zeros) as bf16, one row per
// output pixel — the B operand of the tcgen05 dW GEMM for the stem weight gradient (dW = dY^T patches).
template <typename TIn>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const TIn* __restrict__ frames, int unit_input, int in_h, int in_w, int crop_y, int crop_x, int h, int w, int flip,
                   int oh, int ow, long long total, __nv_bfloat16* __restrict__ patches) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % ow), oy = (int)((idx / ow) % oh);
  const long long f = idx / ((long long)ow * oh);
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const TIn* fbase = frames + (size_t)f * 3 * in_h * in_w;
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    float val = 0.f;
    if (k < 27) {
      const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
      const int y = 2 * oy + ky - 1, x = 2 * ox + kx - 1;
      if (y >= 0 && y < h && x >= 0 && x < w) {
        const int sx = flip ? (w - 1 - x) : x;
        const float raw = (float)fbase[((size_t)ci * in_h + (crop_y + y)) * in_w + (crop_x + sx)];
        val = ((unit_input ? raw : raw / 255.f) - mean[ci]) / stdv[ci];
      }
    }
    v[k] = val;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[8 * q + j];
    store8(patches + idx * 32 + 8 * q, t);
  }
}

}  // namespace tdeed

extern "C" int tdeed_stem_im2col(const void* frames, int frames_dtype, int unit_input, int n_frames, int in_h, int in_w, int crop_y,
                                 int crop_x, int h, int w, int flip, void* patches_bf16, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(frames && patches_bf16 && n_frames > 0 && h > 0 && w > 0 && crop_y >= 0 && crop_x >= 0 && crop_y + h <= in_h &&
                crop_x + w <= in_w, TDEED_ERR_SHAPE, "tdeed_stem_im2col: bad geometry");
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  const long long total = (long long)n_frames * oh * ow;
  const unsigned grid = (unsigned)ceil_div_ll(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (frames_dtype == TDEED_U8)
    stem_im2col_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)frames, unit_input, in_h, in_w, crop_y, crop_x, h, w, flip, oh, ow, total,
                                                       (__nv_bfloat16*)patches_bf16);
  else if (frames_dtype == TDEED_F32)
    stem_im2col_kernel<float><<<grid, 256, 0, st>>>((const float*)frames, unit_input, in_h, in_w, crop_y, crop_x, h, w, flip, oh, ow, total,
                                                     (__nv_bfloat16*)patches_bf16);
  else { set_error("tdeed_stem_im2col: dtype %d", frames_dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_stem_im2col");
}

extern "C" int tdeed_conv3x3g_bwd_data(int dtype, const void* dy, int n, int h, int w, int c, int group_width, int stride,
                                       const float* weight, void* dx, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(dy && weight && dx, TDEED_ERR_SHAPE, "tdeed_conv3x3g_bwd_data: null pointer");
  TDEED_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && c > 0 && c % group_width == 0 && (group_width == 8 || group_width == 16) &&
                (stride == 1 || stride == 2), TDEED_ERR_SHAPE, "tdeed_conv3x3g_bwd_data: bad shape n=%d %dx%dx%d gw=%d s=%d", n, h, w, c, group_width, stride);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16)
    return stride == 1 ? launch_bwd_data<__nv_bfloat16, 1>(dy, n, h, w, c, group_width, weight, dx, st)
                       : launch_bwd_data<__nv_bfloat16, 2>(dy, n, h, w, c, group_width, weight, dx, st);
  if (dtype == TDEED_F32)
    return stride == 1 ? launch_bwd_data<float, 1>(dy, n, h, w, c, group_width, weight, dx, st)
                       : launch_bwd_data<float, 2>(dy, n, h, w, c, group_width, weight, dx, st);
  set_error("tdeed_conv3x3g_bwd_data: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}

namespace tdeed {
// tcgen05 backend (train_conv_tc.cu), bf16 activations
bool conv3x3g_bwd_weight_tc_applicable(int dtype, const void* x, const void* dy, int c);
long long conv3x3g_bwd_weight_tc_workspace_floats(int n, int h, int w, int c, int gw, int stride);
int conv3x3g_bwd_weight_tc_launch(const void* x, const void* dy, int n, int h, int w, int c, int gw, int stride, float* dw, float* ws,
                                  cudaStream_t st);
static bool conv_force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = tdeed::dev_env("TDEED_CONV_BWD_SIMT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
}  // namespace tdeed

extern "C" long long tdeed_conv3x3g_bwd_weight_workspace_floats(int n, int h, int w, int c, int group_width, int stride) {
  using namespace tdeed;
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;
  const long long total = (long long)n * oh * ceil_div(ow, CW_SEG);
  const long long a = (long long)cw_parts(total, c, group_width) * c * group_width * 9;
  const long long b = conv3x3g_bwd_weight_tc_workspace_floats(n, h, w, c, group_width, stride);
  return a > b ? a : b;
}

extern "C" int tdeed_conv3x3g_bwd_weight(int dtype, const void* x, const void* dy, int n, int h, int w, int c, int group_width,
                                         int stride, float* dw, float* workspace, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(x && dy && dw && workspace, TDEED_ERR_SHAPE, "tdeed_conv3x3g_bwd_weight: null pointer");
  TDEED_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % group_width == 0 && (group_width == 8 || group_width == 16) &&
                (stride == 1 || stride == 2), TDEED_ERR_SHAPE, "tdeed_conv3x3g_bwd_weight: bad shape n=%d %dx%dx%d gw=%d s=%d", n, h, w, c, group_width, stride);
  cudaStream_t st = (cudaStream_t)stream;
  if (!conv_force_simt() && conv3x3g_bwd_weight_tc_applicable(dtype, x, dy, c))
    return conv3x3g_bwd_weight_tc_launch(x, dy, n, h, w, c, group_width, stride, dw, workspace, st);
  if (dtype == TDEED_BF16)
    return stride == 1 ? launch_bwd_weight<__nv_bfloat16, 1>(x, dy, n, h, w, c, group_width, dw, workspace, st)
                       : launch_bwd_weight<__nv_bfloat16, 2>(x, dy, n, h, w, c, group_width, dw, workspace, st);
  if (dtype == TDEED_F32)
    return stride == 1 ? launch_bwd_weight<float, 1>(x, dy, n, h, w, c, group_width, dw, workspace, st)
                       : launch_bwd_weight<float, 2>(x, dy, n, h, w, c, group_width, dw, workspace, st);
  set_error("tdeed_conv3x3g_bwd_weight: dtype %d", dtype);
  return TDEED_ERR_UNSUPPORTED;
}

extern "C" long long tdeed_stem_bwd_weight_workspace_floats(void) { return (long long)tdeed::kNumSMs * 4 * 864; }

extern "C" int tdeed_stem_bwd_weight(const void* frames, int frames_dtype, int unit_input, int n_frames, int in_h, int in_w,
                                     int crop_y, int crop_x, int h, int w, int flip, const void* dy, int dy_dtype, float* dw,
                                     float* workspace, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(frames && dy && dw && workspace, TDEED_ERR_SHAPE, "tdeed_stem_bwd_weight: null pointer");
  TDEED_REQUIRE(n_frames > 0 && h > 0 && w > 0 && crop_y >= 0 && crop_x >= 0 && crop_y + h <= in_h && crop_x + w <= in_w,
                TDEED_ERR_SHAPE, "tdeed_stem_bwd_weight: bad geometry");
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  const int tiles_x = ceil_div(ow, SW_TW), tiles_y = ceil_div(oh, SW_TH);
  const long long num_tiles = (long long)n_frames * tiles_x * tiles_y;
  int grid = kNumSMs * 4;
  if (grid > num_tiles) grid = (int)num_tiles;
  cudaStream_t st = (cudaStream_t)stream;
#define SW_CASE(DI, TI, DD, TD) \
  if (frames_dtype == DI && dy_dtype == DD) \
    stem_bwd_weight_kernel<TI, TD><<<grid, SW_THREADS, 0, st>>>((const TI*)frames, unit_input, in_h, in_w, crop_y, crop_x, h, w, flip, \
                                                               (const TD*)dy, oh, ow, tiles_x, tiles_y, num_tiles, workspace); \
  else
  SW_CASE(TDEED_U8, uint8_t, TDEED_BF16, __nv_bfloat16)
  SW_CASE(TDEED_U8, uint8_t, TDEED_F32, float)
  SW_CASE(TDEED_F32, float, TDEED_BF16, __nv_bfloat16)
  SW_CASE(TDEED_F32, float, TDEED_F32, float) {
    set_error("tdeed_stem_bwd_weight: dtypes %d / %d", frames_dtype, dy_dtype);
    return TDEED_ERR_UNSUPPORTED;
  }
#undef SW_CASE
  int rc = check_launch("tdeed_stem_bwd_weight(partial)");
  if (rc) return rc;
  launch_partial_sum(workspace, grid, 864, dw, st);
  return check_launch("tdeed_stem_bwd_weight(final)");
}
