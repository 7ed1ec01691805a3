// (1) preprocessing + stem: u8/f32 planar frames -> crop -> flip -> /255 -> ImageNet normalise ->
// conv3x3 s2 (3->32, BN folded) -> ReLU -> NHWC.  Reference: model/model.py:107,121-129,151-167 + timm stem.
//
// One CTA produces a TILE_H x TILE_W block of output pixels for all 32 channels.  The normalised
// input patch (3 x (2*TILE_H+1) x (2*TILE_W+1)) is staged in shared memory (zero padding applied in
// normalised space, exactly like Conv2d(padding=1) after T.Normalize); weights sit in shared memory
// and are read as warp-wide broadcasts.  Each thread owns one output pixel and 32 fp32 accumulators;
// a warp covers 32 consecutive x so the NHWC store is fully coalesced (32 px * 64 B).
#include "common.cuh"

namespace tdeed {

constexpr int STEM_TW = 32, STEM_TH = 8, STEM_CO = 32;
constexpr int STEM_PW = 2 * STEM_TW + 1, STEM_PH = 2 * STEM_TH + 1;
constexpr int STEM_PWP = STEM_PW + 1;   // padded row pitch

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(STEM_TW * STEM_TH)
stem_kernel(const TIn* __restrict__ frames, int in_h, int in_w, int crop_y, int crop_x, int h, int w, int flip,
            const float* __restrict__ weight, const float* __restrict__ bias, TOut* __restrict__ out,
            int oh, int ow, int relu, int unit_input) {
  __shared__ float s_in[3][STEM_PH][STEM_PWP];
  __shared__ __align__(16) float s_w[27][STEM_CO];   // [ci*9 + ky*3 + kx][co]
  __shared__ float s_b[STEM_CO];
  const int f = blockIdx.z;
  const int oy0 = blockIdx.y * STEM_TH, ox0 = blockIdx.x * STEM_TW;
  const int tid = threadIdx.y * STEM_TW + threadIdx.x;
  const int nthr = STEM_TW * STEM_TH;

  for (int i = tid; i < 27 * STEM_CO; i += nthr) {
    const int co = i % STEM_CO, tap = i / STEM_CO;
    s_w[tap][co] = weight[co * 27 + tap];
  }
  if (tid < STEM_CO) s_b[tid] = bias ? bias[tid] : 0.f;

  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;     // in cropped coordinates
  const TIn* fbase = frames + (size_t)f * 3 * in_h * in_w;
  for (int i = tid; i < 3 * STEM_PH * STEM_PW; i += nthr) {
    const int px = i % STEM_PW, py = (i / STEM_PW) % STEM_PH, ci = i / (STEM_PW * STEM_PH);
    const int y = iy0 + py, x = ix0 + px;
    float v = 0.f;
    if (y >= 0 && y < h && x >= 0 && x < w) {
      const int sx = flip ? (w - 1 - x) : x;
      const float raw = (float)fbase[((size_t)ci * in_h + (crop_y + y)) * in_w + (crop_x + sx)];
      v = ((unit_input ? raw : raw / 255.f) - mean[ci]) / stdv[ci];
    }
    s_in[ci][py][px] = v;
  }
  __syncthreads();

  const int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
  float acc[STEM_CO];
#pragma unroll
  for (int c = 0; c < STEM_CO; ++c) acc[c] = s_b[c];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float v = s_in[ci][2 * threadIdx.y + ky][2 * threadIdx.x + kx];
        const float4* wr = reinterpret_cast<const float4*>(s_w[ci * 9 + ky * 3 + kx]);
#pragma unroll
        for (int q = 0; q < STEM_CO / 4; ++q) {
          const float4 w4 = wr[q];
          acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
        }
      }
  if (oy < oh && ox < ow) {
    TOut* o = out + (((size_t)f * oh + oy) * ow + ox) * STEM_CO;
#pragma unroll
    for (int q = 0; q < STEM_CO / 8; ++q) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = relu ? fmaxf(acc[8 * q + j], 0.f) : acc[8 * q + j];
      store8(o + 8 * q, v);
    }
  }
}

template <typename TIn, typename TOut>
static int launch_stem(const void* frames, int n, int in_h, int in_w, int cy, int cx, int h, int w, int flip,
                       const float* weight, const float* bias, void* out, int relu, int unit_input, cudaStream_t st) {
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  dim3 grid(ceil_div(ow, STEM_TW), ceil_div(oh, STEM_TH), n), block(STEM_TW, STEM_TH);
  stem_kernel<TIn, TOut><<<grid, block, 0, st>>>((const TIn*)frames, in_h, in_w, cy, cx, h, w, flip, weight, bias,
                                                 (TOut*)out, oh, ow, relu, unit_input);
  return check_launch("tdeed_stem_fwd");
}

}  // namespace tdeed

static int stem_dispatch(const char* name, const void* frames, int frames_dtype, int n_frames, int in_h, int in_w,
                         int crop_y, int crop_x, int h, int w, int flip, const float* weight, const float* bias, int relu,
                         int unit_input, void* out, int out_dtype, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(frames && weight && out, TDEED_ERR_SHAPE, "%s: null pointer", name);
  TDEED_REQUIRE(n_frames > 0 && n_frames <= 65535 && h > 0 && w > 0 && crop_y >= 0 && crop_x >= 0 &&
                crop_y + h <= in_h && crop_x + w <= in_w, TDEED_ERR_SHAPE,
                "%s: bad geometry n=%d in=%dx%d crop=(%d,%d) %dx%d", name, n_frames, in_h, in_w, crop_y, crop_x, h, w);
  cudaStream_t st = (cudaStream_t)stream;
#define STEM_CASE(DI, TI, DO, TO) \
  if (frames_dtype == DI && out_dtype == DO) \
    return launch_stem<TI, TO>(frames, n_frames, in_h, in_w, crop_y, crop_x, h, w, flip, weight, bias, out, relu, unit_input, st);
  STEM_CASE(TDEED_U8, uint8_t, TDEED_BF16, __nv_bfloat16)
  STEM_CASE(TDEED_U8, uint8_t, TDEED_F32, float)
  STEM_CASE(TDEED_F32, float, TDEED_BF16, __nv_bfloat16)
  STEM_CASE(TDEED_F32, float, TDEED_F32, float)
#undef STEM_CASE
  set_error("%s: unsupported dtypes %d -> %d", name, frames_dtype, out_dtype);
  return TDEED_ERR_UNSUPPORTED;
}

extern "C" int tdeed_stem_fwd(const void* frames, int frames_dtype, int n_frames, int in_h, int in_w,
                              int crop_y, int crop_x, int h, int w, int flip,
                              const float* weight, const float* bias, void* out, int out_dtype, void* stream) {
  TDEED_REQUIRE(bias, TDEED_ERR_SHAPE, "tdeed_stem_fwd: null pointer");
  return stem_dispatch("tdeed_stem_fwd", frames, frames_dtype, n_frames, in_h, in_w, crop_y, crop_x, h, w, flip, weight, bias, 1, 0,
                       out, out_dtype, stream);
}

// training: raw stem convolution (no BN fold, no ReLU).  unit_input != 0: frames are floats already divided by 255
// (the state the reference's augmentation pipeline leaves them in, model/model.py:107-118).
extern "C" int tdeed_stem_raw_fwd(const void* frames, int frames_dtype, int unit_input, int n_frames, int in_h, int in_w,
                                  int crop_y, int crop_x, int h, int w, int flip, const float* weight, void* out,
                                  int out_dtype, void* stream) {
  return stem_dispatch("tdeed_stem_raw_fwd", frames, frames_dtype, n_frames, in_h, in_w, crop_y, crop_x, h, w, flip, weight, nullptr,
                       0, unit_input, out, out_dtype, stream);
}
