// (1b) bf16 tensor-core stem, optionally fused with the first 1x1 conv of stage 1:
//   frames (u8|f32 planar) -> crop/flip -> /255 -> ImageNet normalise -> conv3x3 s2 (3->32, BN folded) -> ReLU
//   [-> conv1x1 (32 -> n1, BN folded) -> ReLU]                (timm stem + s1.b1.conv1; model/model.py:107,121-129)
//
// The 3x3x3 stem conv is an implicit GEMM with K = 27 (padded to 32): far too thin for a TMA im2col pipeline, but
// a perfect fit for "threads build the operand, tcgen05 does the math":
//   * a CTA tile is 8 x 16 output pixels = 128 GEMM rows = the 128 TMEM lanes;
//   * the normalised input patch (3 x 17 x 33) is staged in shared memory as bf16 (u8 input goes through a
//     256-entry LUT per channel, so normalisation costs one shared-memory read per pixel, no divisions);
//   * every thread packs its pixel's 27 taps into one 64-byte row of the A operand, written straight into the
//     canonical K-major no-swizzle UMMA layout (8x16B core matrices; LBO = 128 B, SBO = 512 B);
//   * one elected thread issues two tcgen05.mma (M=128, N=32, K=16) into TMEM; tcgen05.commit -> mbarrier;
//   * epilogue: tcgen05.ld (lane = pixel) -> bias + ReLU -> bf16.  With the fused conv1 the ReLU'd stem row is
//     written back to shared memory as the A operand of a second MMA pair (N = n1) and only conv1's output
//     (plus the stride-2 subsample of the stem output that the stage-1 shortcut conv needs) goes to HBM —
//     the full-resolution stem activation (the largest tensor of the network) never leaves the SM.
// CTAs are persistent (several per SM) and loop over tiles; phases of different CTAs overlap on an SM.
#include "common.cuh"
#include <climits>

namespace tdeed {

constexpr int ST_TW = 16, ST_TH = 8, ST_PW = 2 * ST_TW + 1, ST_PH = 2 * ST_TH + 1, ST_PWP = ST_PW + 1;
constexpr int ST_THREADS = 128;

struct StemTcParams {
  const void* frames;
  int in_h, in_w, crop_y, crop_x, h, w, flip, oh, ow;
  const __nv_bfloat16* w0;   // [32][32]  (k = ci*9+ky*3+kx, zero padded 27..31)
  const float* b0;           // [32]
  const __nv_bfloat16* w1;   // [n1p][32] or null
  const float* b1;           // [n1]
  int n1, n1p;
  __nv_bfloat16* out_stem;   // NHWC [n, ceil(oh/sub), ceil(ow/sub), 32] or null
  int stem_sub;              // 1: every pixel, 2: even pixels only (input of the stride-2 shortcut conv)
  __nv_bfloat16* out_c1;     // NHWC [n, oh, ow, n1] or null
  int tiles_x, tiles_y, num_tiles;
  uint32_t tmem_cols;
  long long frames_bytes;    // size of the frames tensor (bounds for the aligned word loads)
};

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: 8-row x 16-byte core matrices; next K chunk at +128 B (LBO), next 8-row group at +512 B (SBO)
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void st_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void st_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool st_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(st_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_wait(uint64_t* bar, uint32_t parity) {
  if (st_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!st_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tdeed stem_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// operand row `r` (32 bf16 = 64 B given as 16 packed words) into the canonical no-swizzle layout
__device__ __forceinline__ void st_store_row(uint8_t* tile, int r, const uint32_t (&wds)[16]) {
  uint8_t* base = tile + (r >> 3) * 512 + (r & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc)
    *reinterpret_cast<uint4*>(base + kc * 128) = make_uint4(wds[4 * kc], wds[4 * kc + 1], wds[4 * kc + 2], wds[4 * kc + 3]);
}

// u8 frames: every patch row is 33 contiguous source bytes -> fetch it as <= 10 ALIGNED 32-bit words (4x fewer load
// instructions than byte loads).  Item i = (channel, patch row, word) of thread tid, iteration k.
constexpr int ST_WORDS = 10;                                    // ceil((33 + 3) / 4) + 1
constexpr int ST_ITEMS = 3 * ST_PH * ST_WORDS;
constexpr int ST_WITERS = (ST_ITEMS + ST_THREADS - 1) / ST_THREADS;

__device__ __forceinline__ void stem_load_words(const StemTcParams& p, int tile, int tid, uint32_t (&wv)[ST_WITERS], int (&koff)[ST_WITERS]) {
  const int txi = tile % p.tiles_x, tyi = (tile / p.tiles_x) % p.tiles_y, f = tile / (p.tiles_x * p.tiles_y);
  const int iy0 = 2 * tyi * ST_TH - 1, ix0 = 2 * txi * ST_TW - 1;
  const uint8_t* lo_ok = reinterpret_cast<const uint8_t*>(p.frames);
  const uint8_t* hi_ok = lo_ok + p.frames_bytes;
  const uint8_t* fb = lo_ok + (size_t)f * 3 * p.in_h * p.in_w;
#pragma unroll
  for (int k = 0; k < ST_WITERS; ++k) {
    const int i = tid + k * ST_THREADS;
    const int wq = i % ST_WORDS, py = (i / ST_WORDS) % ST_PH, ci = i / (ST_WORDS * ST_PH);
    const int y = iy0 + py;
    koff[k] = INT_MIN;
    wv[k] = 0u;
    if (i < ST_ITEMS && y >= 0 && y < p.h) {
      // source span of this patch row: offsets 0..32 <-> x = ix0 + (flip ? 32 - o : o)
      const int s_lo = p.flip ? (p.w - 1 - (ix0 + ST_PW - 1)) : ix0;
      const uint8_t* row = fb + ((size_t)ci * p.in_h + (p.crop_y + y)) * p.in_w + p.crop_x + s_lo;
      const int mis = (int)(reinterpret_cast<uintptr_t>(row) & 3);
      const uint8_t* wp = row - mis + 4 * wq;
      koff[k] = 4 * wq - mis;
      if (wp >= lo_ok && wp + 4 <= hi_ok) {
        wv[k] = *reinterpret_cast<const uint32_t*>(wp);
      } else {                                                // first / last word of the whole tensor: byte-wise
        for (int b = 0; b < 4; ++b)
          if (wp + b >= lo_ok && wp + b < hi_ok) wv[k] |= (uint32_t)wp[b] << (8 * b);
      }
    }
  }
}

template <typename TIn>
__global__ void __launch_bounds__(ST_THREADS)
stem_tc_kernel(const StemTcParams p) {
  __shared__ __align__(128) uint8_t sA[128 * 64];          // A operand (im2col rows, later the ReLU'd stem rows)
  __shared__ __align__(128) uint8_t sW0[32 * 64];
  __shared__ __align__(128) uint8_t sW1[64 * 64];
  __shared__ __nv_bfloat16 s_patch[3][ST_PH][ST_PWP];
  __shared__ __nv_bfloat16 s_lut[3][256];
  __shared__ float s_b0[32], s_b1[64];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};

  // ---- one-time setup: weights into the canonical layout, LUT, biases, mbarrier, TMEM ----
  for (int i = tid; i < 32 * 4; i += ST_THREADS) {          // (row, 16-byte chunk)
    const int row = i >> 2, kc = i & 3;
    *reinterpret_cast<uint4*>(sW0 + (row >> 3) * 512 + kc * 128 + (row & 7) * 16) =
        *reinterpret_cast<const uint4*>(p.w0 + row * 32 + kc * 8);
  }
  if (p.w1) {
    for (int i = tid; i < p.n1p * 4; i += ST_THREADS) {
      const int row = i >> 2, kc = i & 3;
      *reinterpret_cast<uint4*>(sW1 + (row >> 3) * 512 + kc * 128 + (row & 7) * 16) =
          *reinterpret_cast<const uint4*>(p.w1 + row * 32 + kc * 8);
    }
    for (int i = tid; i < 64; i += ST_THREADS) s_b1[i] = i < p.n1 ? p.b1[i] : 0.f;
  }
  for (int i = tid; i < 3 * 256; i += ST_THREADS) {
    const int ci = i >> 8, v = i & 255;
    s_lut[ci][v] = __float2bfloat16_rn(((float)v / 255.f - mean[ci]) / stdv[ci]);
  }
  if (tid < 32) s_b0[tid] = p.b0[tid];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(&s_bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(st_smem_u32(&s_tmem)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // weight tiles were written by the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n1p >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;

  const int ty = tid >> 4, tx = tid & 15;
  const TIn* frames = reinterpret_cast<const TIn*>(p.frames);
  const int sub_oh = (p.oh + p.stem_sub - 1) / p.stem_sub, sub_ow = (p.ow + p.stem_sub - 1) / p.stem_sub;

  uint32_t wv[ST_WITERS];
  int koff[ST_WITERS];
  if (sizeof(TIn) == 1 && (int)blockIdx.x < p.num_tiles) stem_load_words(p, blockIdx.x, tid, wv, koff);

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int txi = tile % p.tiles_x, tyi = (tile / p.tiles_x) % p.tiles_y, f = tile / (p.tiles_x * p.tiles_y);
    const int oy0 = tyi * ST_TH, ox0 = txi * ST_TW;
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    const TIn* fbase = frames + (size_t)f * 3 * p.in_h * p.in_w;

    // ---- normalised input patch (zero padding applies in normalised space) ----
    if (sizeof(TIn) == 1) {
      // the words of THIS tile were requested one iteration ago (software pipelining: the global latency of the patch
      // is hidden behind the previous tile's MMAs and epilogue); LUT-normalise them byte by byte
#pragma unroll
      for (int k = 0; k < ST_WITERS; ++k) {
        if (koff[k] == INT_MIN) continue;
        const int i = tid + k * ST_THREADS;
        const int py = (i / ST_WORDS) % ST_PH, ci = i / (ST_WORDS * ST_PH);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int o = koff[k] + b;
          if (o < 0 || o >= ST_PW) continue;
          const int px = p.flip ? (ST_PW - 1 - o) : o;
          const int x = ix0 + px;
          s_patch[ci][py][px] = (x >= 0 && x < p.w) ? s_lut[ci][(wv[k] >> (8 * b)) & 255u] : __float2bfloat16_rn(0.f);
        }
      }
      // rows outside the image are all padding
      for (int i = tid; i < 3 * ST_PH * ST_PW; i += ST_THREADS) {
        const int px = i % ST_PW, py = (i / ST_PW) % ST_PH, ci = i / (ST_PW * ST_PH);
        const int y = iy0 + py;
        if (y < 0 || y >= p.h) s_patch[ci][py][px] = __float2bfloat16_rn(0.f);
      }
    } else {
      // two phases so that all of a thread's global loads are in flight together (the loop is latency bound otherwise)
      constexpr int kElems = 3 * ST_PH * ST_PW;
      constexpr int kIters = (kElems + ST_THREADS - 1) / ST_THREADS;
      TIn raw[kIters];
      bool inside[kIters];
#pragma unroll
      for (int k = 0; k < kIters; ++k) {
        const int i = tid + k * ST_THREADS;
        const int px = i % ST_PW, py = (i / ST_PW) % ST_PH, ci = i / (ST_PW * ST_PH);
        const int y = iy0 + py, x = ix0 + px;
        inside[k] = i < kElems && y >= 0 && y < p.h && x >= 0 && x < p.w;
        raw[k] = TIn(0);
        if (inside[k]) {
          const int sx = p.flip ? (p.w - 1 - x) : x;
          raw[k] = fbase[((size_t)ci * p.in_h + (p.crop_y + y)) * p.in_w + (p.crop_x + sx)];
        }
      }
#pragma unroll
      for (int k = 0; k < kIters; ++k) {
        const int i = tid + k * ST_THREADS;
        if (i >= kElems) continue;
        const int px = i % ST_PW, py = (i / ST_PW) % ST_PH, ci = i / (ST_PW * ST_PH);
        __nv_bfloat16 v = __float2bfloat16_rn(0.f);
        if (inside[k]) v = __float2bfloat16_rn(((float)raw[k] / 255.f - mean[ci]) / stdv[ci]);
        s_patch[ci][py][px] = v;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // previous tile's tcgen05.ld are done
    __syncthreads();

    // ---- im2col row of this thread's pixel -> A operand ----
    {
      uint32_t wds[16];
      const unsigned short* pp = reinterpret_cast<const unsigned short*>(&s_patch[0][0][0]);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t lo = 0, hi = 0;
        const int k0 = 2 * j, k1 = 2 * j + 1;
        if (k0 < 27) lo = pp[((k0 / 9) * ST_PH + 2 * ty + (k0 % 9) / 3) * ST_PWP + 2 * tx + (k0 % 3)];
        if (k1 < 27) hi = pp[((k1 / 9) * ST_PH + 2 * ty + (k1 % 9) / 3) * ST_PWP + 2 * tx + (k1 % 3)];
        wds[j] = lo | (hi << 16);
      }
      st_store_row(sA, tid, wds);
    }
    if (sizeof(TIn) == 1 && tile + (int)gridDim.x < p.num_tiles) stem_load_words(p, tile + gridDim.x, tid, wv, koff);   // prefetch
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a = st_smem_u32(sA), b = st_smem_u32(sW0);
      st_umma(tmem_base, umma_desc_nosw(a), umma_desc_nosw(b), idesc0, 0u);
      st_umma(tmem_base, umma_desc_nosw(a + 256), umma_desc_nosw(b + 256), idesc0, 1u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(st_smem_u32(&s_bar)) : "memory");
    }
    st_wait(&s_bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue 1: bias + ReLU (+ write / + feed the fused 1x1 conv) ----
    const int oy = oy0 + ty, ox = ox0 + tx;
    const bool pix_ok = oy < p.oh && ox < p.ow;
    uint32_t packed[16];
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 16) {
      uint32_t v32[16];
      st_ld16(tmem_lane + (uint32_t)c0, v32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float lo = fmaxf(__uint_as_float(v32[2 * j]) + s_b0[c0 + 2 * j], 0.f);
        const float hi = fmaxf(__uint_as_float(v32[2 * j + 1]) + s_b0[c0 + 2 * j + 1], 0.f);
        packed[c0 / 2 + j] = pack_bf16x2(lo, hi);
      }
    }
    if (p.out_stem && pix_ok && (p.stem_sub == 1 || ((oy & 1) == 0 && (ox & 1) == 0))) {
      __nv_bfloat16* o = p.out_stem + (((size_t)f * sub_oh + oy / p.stem_sub) * sub_ow + ox / p.stem_sub) * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        reinterpret_cast<uint4*>(o)[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    }
    if (p.w1) {
      // the first MMA pair has completed (mbarrier), so sA can be overwritten with the ReLU'd stem rows
      st_store_row(sA, tid, packed);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a = st_smem_u32(sA), b = st_smem_u32(sW1);
        st_umma(tmem_base + 32, umma_desc_nosw(a), umma_desc_nosw(b), idesc1, 0u);
        st_umma(tmem_base + 32, umma_desc_nosw(a + 256), umma_desc_nosw(b + 256), idesc1, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(st_smem_u32(&s_bar)) : "memory");
      }
      st_wait(&s_bar, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __nv_bfloat16* o = p.out_c1 + (((size_t)f * p.oh + oy) * p.ow + ox) * p.n1;
      for (int c0 = 0; c0 < p.n1p; c0 += 16) {
        uint32_t v32[16];
        st_ld16(tmem_lane + 32u + (uint32_t)c0, v32);
        if (!pix_ok) continue;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int n = c0 + 8 * hh;
          if (n >= p.n1) continue;
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = fmaxf(__uint_as_float(v32[8 * hh + q]) + s_b1[n + q], 0.f);
          store8(o + n, v);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace tdeed

extern "C" int tdeed_stem_tc_fwd(const void* frames, int frames_dtype, int n_frames, int in_h, int in_w,
                                 int crop_y, int crop_x, int h, int w, int flip,
                                 const void* w0_bf16, const float* b0, const void* w1_bf16, const float* b1, int n1,
                                 void* out_stem, int stem_sub, void* out_c1, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(frames && w0_bf16 && b0, TDEED_ERR_SHAPE, "tdeed_stem_tc_fwd: null pointer");
  TDEED_REQUIRE(out_stem || out_c1, TDEED_ERR_SHAPE, "tdeed_stem_tc_fwd: no output requested");
  TDEED_REQUIRE(!w1_bf16 || (b1 && out_c1 && n1 > 0 && n1 % 8 == 0 && n1 <= 64), TDEED_ERR_SHAPE,
                "tdeed_stem_tc_fwd: fused conv1 needs bias, output and n1 %% 8 == 0, n1 <= 64 (got %d)", n1);
  TDEED_REQUIRE((w1_bf16 != nullptr) == (out_c1 != nullptr), TDEED_ERR_SHAPE, "tdeed_stem_tc_fwd: out_c1 goes with w1");
  TDEED_REQUIRE(n_frames > 0 && h > 0 && w > 0 && crop_y >= 0 && crop_x >= 0 && crop_y + h <= in_h && crop_x + w <= in_w &&
                (stem_sub == 1 || stem_sub == 2), TDEED_ERR_SHAPE,
                "tdeed_stem_tc_fwd: bad geometry n=%d in=%dx%d crop=(%d,%d) %dx%d sub=%d", n_frames, in_h, in_w, crop_y, crop_x, h, w, stem_sub);
  StemTcParams p{};
  p.frames = frames; p.in_h = in_h; p.in_w = in_w; p.crop_y = crop_y; p.crop_x = crop_x; p.h = h; p.w = w; p.flip = flip;
  p.oh = (h + 1) / 2; p.ow = (w + 1) / 2;
  p.w0 = (const __nv_bfloat16*)w0_bf16; p.b0 = b0; p.w1 = (const __nv_bfloat16*)w1_bf16; p.b1 = b1;
  p.n1 = w1_bf16 ? n1 : 0;
  p.n1p = w1_bf16 ? (n1 + 15) / 16 * 16 : 16;
  p.out_stem = (__nv_bfloat16*)out_stem; p.stem_sub = stem_sub; p.out_c1 = (__nv_bfloat16*)out_c1;
  p.tiles_x = ceil_div(p.ow, ST_TW); p.tiles_y = ceil_div(p.oh, ST_TH);
  const long long nt = (long long)n_frames * p.tiles_x * p.tiles_y;
  TDEED_REQUIRE(nt < (1LL << 31), TDEED_ERR_SHAPE, "tdeed_stem_tc_fwd: too many tiles");
  p.num_tiles = (int)nt;
  p.tmem_cols = (w1_bf16 && p.n1p > 32) ? 128 : 64;
  p.frames_bytes = (long long)n_frames * 3 * in_h * in_w * (frames_dtype == TDEED_U8 ? 1 : 4);
  const int ctas_per_sm = 512 / (int)p.tmem_cols < 6 ? 512 / (int)p.tmem_cols : 6;
  const int grid = (int)(nt < (long long)kNumSMs * ctas_per_sm ? nt : (long long)kNumSMs * ctas_per_sm);
  cudaStream_t st = (cudaStream_t)stream;
  if (frames_dtype == TDEED_U8) stem_tc_kernel<uint8_t><<<grid, ST_THREADS, 0, st>>>(p);
  else if (frames_dtype == TDEED_F32) stem_tc_kernel<float><<<grid, ST_THREADS, 0, st>>>(p);
  else { set_error("tdeed_stem_tc_fwd: frames dtype %d", frames_dtype); return TDEED_ERR_UNSUPPORTED; }
  return check_launch("tdeed_stem_tc_fwd");
}
