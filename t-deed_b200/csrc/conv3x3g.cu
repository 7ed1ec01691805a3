// (3) grouped 3x3 convolution + folded BN + ReLU (timm Bottleneck.conv2; group width 8 or 16).
//
// Work item = (output row, strip of PX=4 output pixels, output-channel octet).  A thread keeps
// 4 x 8 fp32 accumulators and, per input row, the strip's input pixels (8 channels each) in
// registers, so every shared-memory weight fetch (two float4 = 8 output channels) feeds 32 FMAs.
// Weights of the CTA's octets are staged (transposed to [ky][kx][ci][co8]) in shared memory with a
// per-octet pitch == 4 (mod 32) words so that the float4 reads of a quarter-warp hit distinct banks.
// Adjacent threads own adjacent octets of the same pixel strip -> NHWC loads/stores are coalesced.
#include "common.cuh"

namespace tdeed {

constexpr int C3_PX = 4;
constexpr int C3_THREADS = 256;
constexpr int C3_MAX_UNITS = 16;   // output octets per CTA

template <typename T, int STRIDE>
__global__ void __launch_bounds__(C3_THREADS)
conv3x3g_kernel(const T* __restrict__ in, int h, int w, int c, int gw, const float* __restrict__ weight,
                const float* __restrict__ bias, T* __restrict__ out, int oh, int ow, int units_per_cta, int relu) {
  extern __shared__ __align__(16) float s_w[];
  const int n_units = c / 8;
  const int u0 = blockIdx.y * units_per_cta;
  const int ucnt = min(units_per_cta, n_units - u0);
  const int per_unit = 9 * gw * 8;
  const int pitch = per_unit + 4;
  const int f = blockIdx.z;

  // stage weights: s_w[ul][ky][kx][ci][co8] = weight[(8*(u0+ul)+co8)][ci][ky][kx]
  for (int i = threadIdx.x; i < ucnt * per_unit; i += C3_THREADS) {
    const int ul = i / per_unit, r = i - ul * per_unit;
    const int co8 = r & 7, ci = (r >> 3) % gw, tap = r / (8 * gw);
    s_w[ul * pitch + r] = weight[((size_t)(8 * (u0 + ul) + co8) * gw + ci) * 9 + tap];
  }
  __syncthreads();

  const int strips = (ow + C3_PX - 1) / C3_PX;
  const int item = blockIdx.x * C3_THREADS + threadIdx.x;
  if (item >= oh * strips * ucnt) return;
  const int ul = item % ucnt;
  const int strip = (item / ucnt) % strips;
  const int oy = item / (ucnt * strips);
  const int u = u0 + ul;
  const int cin0 = (u * 8 / gw) * gw;      // first input channel of this octet's group
  const int ox0 = strip * C3_PX;
  constexpr int NIN = (C3_PX - 1) * STRIDE + 3;

  float acc[C3_PX][8];
#pragma unroll
  for (int p = 0; p < C3_PX; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;

  const T* fin = in + (size_t)f * h * w * c;
  const float* wu = s_w + ul * pitch;
  for (int cio = 0; cio < gw; cio += 8) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * STRIDE + ky - 1;
      if (iy < 0 || iy >= h) continue;
      float v[NIN][8];
#pragma unroll
      for (int x = 0; x < NIN; ++x) {
        const int ix = ox0 * STRIDE + x - 1;
        if (ix >= 0 && ix < w) {
          load8(fin + ((size_t)iy * w + ix) * c + cin0 + cio, v[x]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[x][j] = 0.f;
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          const float4* wp = reinterpret_cast<const float4*>(wu + ((ky * 3 + kx) * gw + cio + ci) * 8);
          const float4 wa = wp[0], wb = wp[1];
#pragma unroll
          for (int p = 0; p < C3_PX; ++p) {
            const float a = v[p * STRIDE + kx][ci];
            acc[p][0] = fmaf(a, wa.x, acc[p][0]);
            acc[p][1] = fmaf(a, wa.y, acc[p][1]);
            acc[p][2] = fmaf(a, wa.z, acc[p][2]);
            acc[p][3] = fmaf(a, wa.w, acc[p][3]);
            acc[p][4] = fmaf(a, wb.x, acc[p][4]);
            acc[p][5] = fmaf(a, wb.y, acc[p][5]);
            acc[p][6] = fmaf(a, wb.z, acc[p][6]);
            acc[p][7] = fmaf(a, wb.w, acc[p][7]);
          }
        }
      }
    }
  }

  float b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = bias ? bias[u * 8 + j] : 0.f;
  T* fout = out + (size_t)f * oh * ow * c;
#pragma unroll
  for (int p = 0; p < C3_PX; ++p) {
    const int ox = ox0 + p;
    if (ox >= ow) break;
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = relu ? fmaxf(acc[p][j] + b[j], 0.f) : acc[p][j] + b[j];
    store8(fout + ((size_t)oy * ow + ox) * c + u * 8, r);
  }
}

template <typename T, int STRIDE>
static int launch_conv3(const void* in, int n, int h, int w, int c, int gw, const float* weight, const float* bias,
                        void* out, int relu, cudaStream_t st) {
  const int oh = (h + STRIDE - 1) / STRIDE, ow = (w + STRIDE - 1) / STRIDE;
  const int n_units = c / 8;
  const int upc = n_units < C3_MAX_UNITS ? n_units : C3_MAX_UNITS;
  const int strips = (ow + C3_PX - 1) / C3_PX;
  const size_t smem = (size_t)upc * (9 * gw * 8 + 4) * sizeof(float);
  auto kern = conv3x3g_kernel<T, STRIDE>;
  static size_t smem_set = 48 * 1024;   // per template instantiation
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "conv3x3g: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = smem;
  }
  dim3 grid(ceil_div(oh * strips * upc, C3_THREADS), ceil_div(n_units, upc), n);
  kern<<<grid, C3_THREADS, smem, st>>>((const T*)in, h, w, c, gw, weight, bias, (T*)out, oh, ow, upc, relu);
  return check_launch("tdeed_conv3x3g_fwd");
}

}  // namespace tdeed

static int conv3x3g_dispatch(const char* name, int dtype, const void* in, int n, int h, int w, int c, int group_width, int stride,
                             const float* weight, const float* bias, int relu, void* out, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(in && weight && out, TDEED_ERR_SHAPE, "%s: null pointer", name);
  TDEED_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && c > 0 && c % group_width == 0, TDEED_ERR_SHAPE,
                "%s: bad shape n=%d %dx%dx%d gw=%d", name, n, h, w, c, group_width);
  TDEED_REQUIRE(group_width == 8 || group_width == 16, TDEED_ERR_UNSUPPORTED,
                "%s: group width %d (RegNetY-200MF/800MF use 8/16)", name, group_width);
  TDEED_REQUIRE(stride == 1 || stride == 2, TDEED_ERR_UNSUPPORTED, "%s: stride %d", name, stride);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TDEED_BF16) {
    return stride == 1 ? launch_conv3<__nv_bfloat16, 1>(in, n, h, w, c, group_width, weight, bias, out, relu, st)
                       : launch_conv3<__nv_bfloat16, 2>(in, n, h, w, c, group_width, weight, bias, out, relu, st);
  } else if (dtype == TDEED_F32) {
    return stride == 1 ? launch_conv3<float, 1>(in, n, h, w, c, group_width, weight, bias, out, relu, st)
                       : launch_conv3<float, 2>(in, n, h, w, c, group_width, weight, bias, out, relu, st);
  }
  set_error("%s: dtype %d", name, dtype);
  return TDEED_ERR_UNSUPPORTED;
}

extern "C" int tdeed_conv3x3g_fwd(int dtype, const void* in, int n, int h, int w, int c, int group_width, int stride,
                                  const float* weight, const float* bias, void* out, void* stream) {
  TDEED_REQUIRE(bias, TDEED_ERR_SHAPE, "tdeed_conv3x3g_fwd: null pointer");
  return conv3x3g_dispatch("tdeed_conv3x3g_fwd", dtype, in, n, h, w, c, group_width, stride, weight, bias, 1, out, stream);
}

// training: the raw convolution (no bias, no activation); BatchNorm statistics are taken from this output
extern "C" int tdeed_conv3x3g_raw_fwd(int dtype, const void* in, int n, int h, int w, int c, int group_width, int stride,
                                      const float* weight, void* out, void* stream) {
  return conv3x3g_dispatch("tdeed_conv3x3g_raw_fwd", dtype, in, n, h, w, c, group_width, stride, weight, nullptr, 0, out, stream);
}
