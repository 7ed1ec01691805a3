// Training-time clip augmentation (model/model.py:77-84,154-157 of the reference: torchvision ColorJitter hue / saturation /
// brightness / contrast, GaussianBlur(5), RandomHorizontalFlip, each RandomApply'd per clip) as three fused passes instead of
// the ~60 elementwise torch kernels the torchvision tensor ops launch per clip:
//   tdeed_aug_color          crop + scale (x/255) + hue -> saturation -> brightness (any subset), planar fp32 out in [0, 1]
//   tdeed_aug_gray_mean      per-frame mean of the grayscale image (the contrast pivot), fixed-order reduction
//   tdeed_aug_contrast_blur_flip   contrast (pointwise, pivot per frame) -> 5x5 separable Gaussian with reflect padding -> flip
// The arithmetic follows torchvision.transforms._functional_tensor (rgb_to_grayscale weights 0.2989/0.587/0.114, _rgb2hsv /
// _hsv2rgb, _blend with clamp to [0, 1], _get_gaussian_kernel1d) so that results agree to float rounding; the random
// parameters are drawn on the host by tdeed_b200/augment.py with torchvision's distributions.
#include "common.cuh"

namespace tdeed {

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
__device__ __forceinline__ float gray_of(float r, float g, float b) { return 0.2989f * r + 0.587f * g + 0.114f * b; }

__device__ inline void hue_shift(float& r, float& g, float& b, float hue) {
  // _rgb2hsv
  const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
  const bool eqc = maxc == minc;
  const float cr = maxc - minc;
  const float s = cr / (eqc ? 1.f : maxc);
  const float div = eqc ? 1.f : cr;
  const float rc = (maxc - r) / div, gc = (maxc - g) / div, bc = (maxc - b) / div;
  float h = 0.f;
  if (maxc == r) h = bc - gc;
  else if (maxc == g) h = 2.f + rc - bc;
  else h = 4.f + gc - rc;
  h = fmodf(h / 6.f + 1.f, 1.f);
  // shift: (h + hue_factor) % 1.0 with Python/torch remainder semantics (result in [0, 1))
  h = h + hue;
  h = h - floorf(h);
  // _hsv2rgb
  const float v = maxc;
  const float h6 = h * 6.f;
  const float fi = floorf(h6);
  const float f = h6 - fi;
  int i = (int)fi;
  i = ((i % 6) + 6) % 6;
  const float p = clamp01(v * (1.f - s));
  const float q = clamp01(v * (1.f - f * s));
  const float t = clamp01(v * (1.f - s * (1.f - f)));
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(256)
aug_color_kernel(const TIn* __restrict__ in, float in_scale, int in_h, int in_w, int crop_y, int crop_x, int h, int w, long long total,
                 int hue_on, float hue, int sat_on, float sat, int bri_on, float bri, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % w), y = (int)((idx / w) % h);
  const long long f = idx / ((long long)w * h);
  const size_t plane_in = (size_t)in_h * in_w, plane_out = (size_t)h * w;
  const TIn* src = in + (size_t)f * 3 * plane_in + (size_t)(crop_y + y) * in_w + crop_x + x;
  float r = (float)src[0] * in_scale, g = (float)src[plane_in] * in_scale, b = (float)src[2 * plane_in] * in_scale;
  if (hue_on) hue_shift(r, g, b, hue);
  if (sat_on) {                      // _blend(img, gray, f) = clamp(f*img + (1-f)*gray)
    const float gr = gray_of(r, g, b);
    r = clamp01(sat * r + (1.f - sat) * gr);
    g = clamp01(sat * g + (1.f - sat) * gr);
    b = clamp01(sat * b + (1.f - sat) * gr);
  }
  if (bri_on) {                      // _blend(img, 0, f)
    r = clamp01(bri * r);
    g = clamp01(bri * g);
    b = clamp01(bri * b);
  }
  float* dst = out + (size_t)f * 3 * plane_out + (size_t)y * w + x;
  dst[0] = r;
  dst[plane_out] = g;
  dst[2 * plane_out] = b;
}

// CTA per frame: mean over all pixels of the grayscale image
__global__ void __launch_bounds__(256) aug_gray_mean_kernel(const float* __restrict__ x, int hw, float* __restrict__ mean) {
  __shared__ float s_red[32];
  const float* p = x + (size_t)blockIdx.x * 3 * hw;
  float s = 0.f;
  for (int i = threadIdx.x; i < hw; i += 256) s += gray_of(p[i], p[hw + i], p[2 * hw + i]);
  s = block_sum(s, s_red);
  if (threadIdx.x == 0) mean[blockIdx.x] = s / (float)hw;
}

struct BlurK { float k[5]; };

__global__ void __launch_bounds__(256)
aug_cbf_kernel(const float* __restrict__ x, int h, int w, long long total, int con_on, float con, const float* __restrict__ mean,
               int blur_on, BlurK kern, int flip, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;                        // total = frames * 3 * h * w
  const int xo = (int)(idx % w), y = (int)((idx / w) % h);
  const long long plane = idx / ((long long)w * h);          // frame*3 + channel
  const long long f = plane / 3;
  const float* src = x + (size_t)plane * h * w;
  const float piv = con_on ? (1.f - con) * mean[f] : 0.f;
  const int xs = flip ? (w - 1 - xo) : xo;                    // source column of this output pixel
  float v;
  if (blur_on) {
    float acc = 0.f;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {
      int yy = y + dy;
      yy = yy < 0 ? -yy : (yy >= h ? 2 * h - 2 - yy : yy);   // reflect (no edge repeat), like F.pad(mode='reflect')
      float row = 0.f;
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        int xx = xs + dx;
        xx = xx < 0 ? -xx : (xx >= w ? 2 * w - 2 - xx : xx);
        float t = src[(size_t)yy * w + xx];
        if (con_on) t = clamp01(con * t + piv);
        row = fmaf(kern.k[dx + 2], t, row);
      }
      acc = fmaf(kern.k[dy + 2], row, acc);
    }
    v = acc;
  } else {
    v = src[(size_t)y * w + xs];
    if (con_on) v = clamp01(con * v + piv);
  }
  out[idx] = v;
}

// mixup of two uint8 clips per sample (model/model.py:228-254 of the reference): out = fl32(l) * a + fl32(1 - l) * b with the
// reference's three roundings (two products, one sum); 16 pixels per thread (two 16-byte loads, four 16-byte stores)
__global__ void __launch_bounds__(256)
mixup_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, const float* __restrict__ lam, long long per_sample16,
                long long total16, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total16) return;
  const long long smp = i / per_sample16;
  const float la = lam[2 * smp], lb = lam[2 * smp + 1];
  const uint4 ra = reinterpret_cast<const uint4*>(a)[i], rb = reinterpret_cast<const uint4*>(b)[i];
  const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
  float4* o = reinterpret_cast<float4*>(out) + i * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      v[j] = __fadd_rn(__fmul_rn(la, (float)((wa[q] >> (8 * j)) & 0xffu)), __fmul_rn(lb, (float)((wb[q] >> (8 * j)) & 0xffu)));
    o[q] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace tdeed

using namespace tdeed;

extern "C" int tdeed_mixup_u8(const void* a, const void* b, const float* lam, int n_samples, long long per_sample, float* out,
                              void* stream) {
  TDEED_REQUIRE(a && b && lam && out && n_samples > 0 && per_sample > 0 && per_sample % 16 == 0, TDEED_ERR_SHAPE,
                "tdeed_mixup_u8: n=%d per_sample=%lld (must be a multiple of 16)", n_samples, per_sample);
  TDEED_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                TDEED_ERR_SHAPE, "tdeed_mixup_u8: pointers must be 16-byte aligned");
  const long long total16 = (long long)n_samples * (per_sample / 16);
  mixup_u8_kernel<<<(unsigned)ceil_div_ll(total16, 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)a, (const uint8_t*)b, lam,
                                                                                         per_sample / 16, total16, out);
  return check_launch("tdeed_mixup_u8");
}

extern "C" int tdeed_aug_color(const void* frames, int frames_dtype, float in_scale, int n_frames, int in_h, int in_w, int crop_y,
                               int crop_x, int h, int w, int hue_on, float hue, int sat_on, float sat, int bri_on, float bri,
                               float* out, void* stream) {
  TDEED_REQUIRE(frames && out && n_frames > 0 && h > 0 && w > 0 && crop_y >= 0 && crop_x >= 0 && crop_y + h <= in_h && crop_x + w <= in_w,
                TDEED_ERR_SHAPE, "tdeed_aug_color: bad geometry");
  const long long total = (long long)n_frames * h * w;
  const unsigned grid = (unsigned)ceil_div_ll(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (frames_dtype == TDEED_U8)
    aug_color_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)frames, in_scale, in_h, in_w, crop_y, crop_x, h, w, total, hue_on, hue,
                                                     sat_on, sat, bri_on, bri, out);
  else if (frames_dtype == TDEED_F32)
    aug_color_kernel<float><<<grid, 256, 0, st>>>((const float*)frames, in_scale, in_h, in_w, crop_y, crop_x, h, w, total, hue_on, hue,
                                                   sat_on, sat, bri_on, bri, out);
  else { set_error("tdeed_aug_color: dtype %d", frames_dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_aug_color");
}

extern "C" int tdeed_aug_gray_mean(const float* x, int n_frames, int hw, float* mean, void* stream) {
  TDEED_REQUIRE(x && mean && n_frames > 0 && hw > 0, TDEED_ERR_SHAPE, "tdeed_aug_gray_mean: bad arguments");
  aug_gray_mean_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(x, hw, mean);
  return check_launch("tdeed_aug_gray_mean");
}

extern "C" int tdeed_aug_contrast_blur_flip(const float* x, int n_frames, int h, int w, int con_on, float con, const float* mean,
                                            int blur_on, const float* kernel1d_host, int flip, float* out, void* stream) {
  TDEED_REQUIRE(x && out && x != out && n_frames > 0 && h > 2 && w > 2 && (!con_on || mean) && (!blur_on || kernel1d_host), TDEED_ERR_SHAPE,
                "tdeed_aug_contrast_blur_flip: bad arguments");
  BlurK k{};
  if (blur_on)
    for (int i = 0; i < 5; ++i) k.k[i] = kernel1d_host[i];
  const long long total = (long long)n_frames * 3 * h * w;
  aug_cbf_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(x, h, w, total, con_on, con, mean, blur_on, k, flip, out);
  return check_launch("tdeed_aug_contrast_blur_flip");
}
