// (1c) Stem as a shifted-descriptor implicit GEMM on raw pixels (uint8 frames, bf16 path), fused with s1.b1.conv1.
//
// The first tcgen05 stem (stem_tc.cu) was INSTRUCTION bound (ncu r1h: 6 900 warp-instructions per 128-pixel tile, 70 % issue
// utilisation): every input byte went through a LUT normalisation + 16-bit shared-memory store, and every output pixel
// rebuilt its 27-tap im2col row from shared memory.  This version removes both:
//   * normalisation is folded into the weights: conv(((x/255) - mean)/std) = sum (w/(255 std)) * x_raw - sum w*mean/std, and
//     zero padding in normalised space = padding with the raw value 255*mean, so the A operand holds RAW pixel values
//     (0..255 are exact in bf16; no LUT, three integer->bf16 conversions per pixel);
//   * no im2col: like conv3x3g_tc.cu, the window is staged once per tile group as [parity plane][position][16 B] chunks (one
//     chunk = the 3 channels of a pixel + 5 zeros) in a padded position space, and the taps are descriptor start offsets.
//     K = 16 per MMA = TWO taps: the second K-chunk of the A descriptor is the same plane one tap further (LBO = distance
//     between the two taps in bytes), so the 9 taps need 5 MMAs (M=128, N=32, K=16) per 128 output positions.
// Warp roles (288 threads): warps 0-3 producers (position table, byte loads, chunk stores), warp 4 MMA issuer (2 M-tiles per
// group into one of two TMEM accumulator sets), warps 5-8 epilogue: bias + ReLU, stride-2 subsample store (the input of the
// s1.b1 shortcut conv), then the ReLU'd rows go back to shared memory as the A operand of the fused 32 -> n1 1x1 conv
// (2 MMAs, K = 32), whose accumulator is stored as the full-resolution output.  The 32-channel stem activation at full
// resolution never reaches HBM.
#include "common.cuh"
#include <cstring>

namespace tdeed {

constexpr int S2_THREADS = 32 * 9;
constexpr int S2_MT = 2;                       // 128-position M tiles per group
constexpr int S2_GROUP = S2_MT * 128;
constexpr int S2_NPAIR = 5;
constexpr int S2_NBUF = 2;                     // window buffers per CTA (staging of group i+1 overlaps the MMAs / epilogue of group i)

struct Stem2Params {
  const uint8_t* frames;
  int n, in_h, in_w, crop_y, crop_x, H, W, flip, Ho, Wo;
  int GH, GW, G;
  long long total_pos;
  int ngroups;
  int min_off, npos, npos_pad;
  int pair_plane[S2_NPAIR], pair_off[S2_NPAIR], pair_lbo[S2_NPAIR];   // offsets / LBO in positions
  const uint8_t* wimg;       // [5][1024 B] B tiles [32 out][16 k] canonical K-major no-swizzle
  const float* b0;           // [32] folded bias (BN shift - sum w*mean/std)
  long long frames_bytes;    // size of the frames tensor (bounds of the aligned word loads)
  uint32_t pad_rg, pad_b;    // raw padding pixel (255*mean) as packed bf16: (g << 16 | r), (0 << 16 | b)
  const __nv_bfloat16* w1;   // [n1p][32]
  const float* b1;
  int n1, n1p;
  __nv_bfloat16* out_stem;   // [n, sub_oh, sub_ow, 32] (every stem_sub-th pixel) or null
  int stem_sub, sub_oh, sub_ow;
  __nv_bfloat16* out_c1;     // [n, Ho, Wo, n1]
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t s2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t s2_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void s2_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void s2_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool s2_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"     // suspend-time hint: sleep in hardware instead of
      "selp.u32 %0, 1, 0, p;\n\t}"                                            // spinning (ncu r1h: half of all issued instructions
      : "=r"(ok) : "r"(s2_smem_u32(bar)), "r"(parity), "r"(4000u) : "memory");   // were wait-loop branches stealing producer slots)
  return ok != 0;
}
__device__ __forceinline__ void s2_wait(uint64_t* bar, uint32_t parity) {
  if (s2_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!s2_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tdeed stem_tc2: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void s2_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void s2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s2_smem_u32(bar)) : "memory");
}
// bf16 bit pattern of an integer 0..255 (exact): float(2^23 + b) - 2^23, top 16 bits
__device__ __forceinline__ uint32_t s2_u8_bf16(uint32_t b) {
  return __float_as_uint(__uint_as_float(0x4B000000u | b) - 8388608.f) >> 16;
}

__global__ void __launch_bounds__(S2_THREADS, 3)      // <= 75 registers: 3 CTAs / SM (the kernel is a latency chain per CTA)
stem_tc2_kernel(const Stem2Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t in_bytes = (size_t)4 * p.npos_pad * 16;
  uint8_t* sW = smem;                                        // [5][1024]
  uint8_t* sW1 = sW + S2_NPAIR * 1024;                       // [n1p/8][4][8][16 B] = n1p * 64 B  (<= 4096)
  uint8_t* sA2 = sW1 + 4096;                                 // [M tile][16 row groups][4 k chunks][8][16 B] = S2_MT x 8192
  uint8_t* sIn = sA2 + S2_MT * 8192;                         // [S2_NBUF][4 planes][npos_pad][16]
  float* s_b0 = reinterpret_cast<float*>(sIn + S2_NBUF * in_bytes);
  float* s_b1 = s_b0 + 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b1 + 64);
  uint64_t* full_bar = bars;          // [2] window staged        (128 producer arrivals)
  uint64_t* empty_bar = bars + 2;     // [2] window consumed      (tcgen05.commit)
  uint64_t* tfull_bar = bars + 4;     // [2] accumulators ready   (tcgen05.commit)
  uint64_t* tempty_bar = bars + 6;    // [2] accumulators drained (128 epilogue arrivals)
  uint64_t* c1_bar = bars + 8;        // fused conv1 MMA done     (tcgen05.commit)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 9);

  for (int i = tid; i < S2_NPAIR * 64; i += S2_THREADS) reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  for (int i = tid; i < p.n1p * 4; i += S2_THREADS) {
    const int row = i >> 2, kc = i & 3;
    *reinterpret_cast<uint4*>(sW1 + (row >> 3) * 512 + kc * 128 + (row & 7) * 16) = *reinterpret_cast<const uint4*>(p.w1 + row * 32 + kc * 8);
  }
  for (int i = tid; i < 32; i += S2_THREADS) s_b0[i] = p.b0[i];
  for (int i = tid; i < 64; i += S2_THREADS) s_b1[i] = i < p.n1 ? p.b1[i] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2_smem_u32(&full_bar[i])), "r"(128));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2_smem_u32(&empty_bar[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2_smem_u32(&tfull_bar[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2_smem_u32(&tempty_bar[i])), "r"(128));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2_smem_u32(c1_bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2_smem_u32(s_tmem)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  // TMEM columns: stem accumulator of M tile m at m*32 (32 channels); fused conv1 accumulator at 64
  const uint32_t acc2_col = 64u;

  if (warp < 4) {
    // ===== producers (128 threads) =====
    // item = (4 consecutive positions, row parity py): the 8 source bytes per channel that feed both column-parity planes
    // are fetched as three aligned 32-bit words (funnel-shifted), 9 loads for 8 chunks; positions that straddle a grid row,
    // touch the padding or the tensor's ends take the byte-wise path.
    const size_t plane_in = (size_t)p.in_h * p.in_w;
    const uint8_t* lo_ok = p.frames;
    const uint8_t* hi_ok = p.frames + p.frames_bytes;
    const int nquads = (p.npos + 3) >> 2;
    uint32_t it = 0;
    for (int grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x, ++it) {
      const uint32_t buf = it % S2_NBUF;
      s2_wait(&empty_bar[buf], ((it / S2_NBUF) & 1u) ^ 1u);
      const long long q_lo = (long long)grp * S2_GROUP + p.min_off;
      uint8_t* dstb = sIn + buf * in_bytes;
      for (int e = tid; e < nquads * 2; e += 128) {
        const int py = e & 1, s0 = (e >> 1) << 2;
        const long long L0 = q_lo + s0;
        int f = 0, U = 0, V = 0;
        bool fast = false;
        if (L0 >= 0 && L0 + 3 < p.total_pos) {
          const uint32_t Lu = (uint32_t)L0;
          f = (int)(Lu / (uint32_t)p.G);
          const uint32_t rem = Lu - (uint32_t)f * (uint32_t)p.G;
          U = (int)(rem / (uint32_t)p.GW);
          V = (int)(rem - (uint32_t)U * (uint32_t)p.GW);
          const int iy = 2 * (U - 1) + py, ix0 = 2 * (V - 1);
          if (U >= 1 && V >= 1 && V + 3 < p.GW && iy < p.H && ix0 + 7 < p.W && s0 + 3 < p.npos) {
            const int a0 = p.flip ? (p.W - 8 - ix0) : ix0;
            const uint8_t* row = p.frames + (size_t)f * 3 * plane_in + (size_t)(p.crop_y + iy) * p.in_w + p.crop_x + a0;
            const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(row) & 3);
            const uint8_t* al = row - mis;
            if (al >= lo_ok && al + 2 * plane_in + 12 <= hi_ok) {
              fast = true;
              uint32_t lo[3], hi[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(al + (size_t)c * plane_in);
                const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
                lo[c] = __funnelshift_r(w0, w1, 8 * mis);
                hi[c] = __funnelshift_r(w1, w2, 8 * mis);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {                 // j = offset of the pixel inside the 8-pixel window (x order)
                const int jb = p.flip ? 7 - j : j;           // byte inside the fetched window
                uint32_t ch[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) ch[c] = s2_u8_bf16(((jb < 4 ? lo[c] : hi[c]) >> (8 * (jb & 3))) & 255u);
                const int pl = py * 2 + (j & 1);
                *reinterpret_cast<uint4*>(dstb + ((size_t)pl * p.npos_pad + s0 + (j >> 1)) * 16) = make_uint4(ch[0] | (ch[1] << 16), ch[2], 0u, 0u);
              }
            }
          }
        }
        if (!fast) {
          // byte-wise path (quads that straddle a grid row, touch the padding or the tensor's ends: ~8 % of the items, but
          // almost every warp holds one).  Addresses first, then ALL loads of two positions (12 bytes) in flight at once from
          // always-valid addresses, then pack + store: the former one-pixel-at-a-time loop exposed eight dependent global
          // latencies per item and dominated the producers' time (ncu r1h: 3x the samples of the word path).
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const uint8_t* src[4];
            bool okp[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int s = s0 + h2 * 2 + (u >> 1), px = u & 1;
              const long long L = q_lo + s;
              bool ok = false;
              const uint8_t* sp = p.frames;
              if (s < p.npos && L >= 0 && L < p.total_pos) {
                const uint32_t Lu = (uint32_t)L;
                const int ff = (int)(Lu / (uint32_t)p.G);
                const uint32_t rem = Lu - (uint32_t)ff * (uint32_t)p.G;
                const int UU = (int)(rem / (uint32_t)p.GW), VV = (int)(rem - (uint32_t)UU * (uint32_t)p.GW);
                const int iy = 2 * (UU - 1) + py, ix = 2 * (VV - 1) + px;
                if (UU >= 1 && VV >= 1 && iy < p.H && ix < p.W) {
                  const int sx = p.flip ? (p.W - 1 - ix) : ix;
                  sp = p.frames + (size_t)ff * 3 * plane_in + (size_t)(p.crop_y + iy) * p.in_w + p.crop_x + sx;
                  ok = true;
                }
              }
              src[u] = sp;
              okp[u] = ok;
            }
            uint32_t cr[4], cg[4], cb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {            // p.frames + {0, plane, 2 plane} is always inside the tensor
              cr[u] = __ldg(src[u]);
              cg[u] = __ldg(src[u] + plane_in);
              cb[u] = __ldg(src[u] + 2 * plane_in);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int s = s0 + h2 * 2 + (u >> 1), px = u & 1;
              if (s >= p.npos) continue;
              const uint32_t rg = okp[u] ? (s2_u8_bf16(cr[u]) | (s2_u8_bf16(cg[u]) << 16)) : p.pad_rg;
              const uint32_t bb = okp[u] ? s2_u8_bf16(cb[u]) : p.pad_b;
              *reinterpret_cast<uint4*>(dstb + ((size_t)(py * 2 + px) * p.npos_pad + s) * 16) = make_uint4(rg, bb, 0u, 0u);
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
      s2_arrive(&full_bar[buf]);
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t b0 = s2_smem_u32(sW);
    uint32_t it = 0;
    for (int grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x, ++it) {
      const uint32_t buf = it % S2_NBUF;
      s2_wait(&tempty_bar[0], (it & 1u) ^ 1u);           // ONE accumulator set (64 TMEM columns)
      s2_wait(&full_bar[buf], (it / S2_NBUF) & 1u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a0 = s2_smem_u32(sIn + buf * in_bytes);
#pragma unroll
        for (int m = 0; m < S2_MT; ++m) {
#pragma unroll
          for (int q = 0; q < S2_NPAIR; ++q) {
            const uint32_t a = a0 + ((uint32_t)p.pair_plane[q] * (uint32_t)p.npos_pad + (uint32_t)(m * 128 + p.pair_off[q] - p.min_off)) * 16u;
            // A: rows 16 B apart (8-row groups 128 B apart = SBO); second K chunk = the partner tap, pair_lbo positions further
            s2_umma(tmem_base + (uint32_t)m * 32u, s2_desc(a, (uint32_t)p.pair_lbo[q] * 16u, 128u),
                    s2_desc(b0 + (uint32_t)q * 1024u, 128u, 256u), idesc, q != 0 ? 1u : 0u);
          }
        }
        s2_commit(&empty_bar[buf]);
        s2_commit(&tfull_bar[0]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue (128 threads, thread = output position) =====
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n1p >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t it = 0, c1_phase = 0;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(lg * 32) << 16);
    for (int grp = blockIdx.x; grp < p.ngroups; grp += gridDim.x, ++it) {
      s2_wait(&tfull_bar[0], it & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // drain BOTH stem accumulators into registers first, so that the next group's MMAs overlap the conv1 chains below
      uint32_t packed[S2_MT][16];
#pragma unroll
      for (int m = 0; m < S2_MT; ++m) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v32[16];
          s2_ld16(tmem_lane + (uint32_t)m * 32u + (uint32_t)half * 16u, v32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = fmaxf(__uint_as_float(v32[2 * j]) + s_b0[half * 16 + 2 * j], 0.f);
            const float b = fmaxf(__uint_as_float(v32[2 * j + 1]) + s_b0[half * 16 + 2 * j + 1], 0.f);
            packed[m][half * 8 + j] = pack_bf16x2(a, b);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      s2_arrive(&tempty_bar[0]);
      // fused s1.b1.conv1 for BOTH M tiles in one round: the ReLU'd rows are the A operands (K = 32) of 2 x 2 MMAs into two
      // conv1 accumulators — one shared-memory hand-off / barrier / commit / wait per group instead of one per tile
      bool okm[S2_MT];
      int fm[S2_MT], oym[S2_MT], oxm[S2_MT];
#pragma unroll
      for (int m = 0; m < S2_MT; ++m) {
        const long long L = (long long)grp * S2_GROUP + m * 128 + r;
        bool ok = L < p.total_pos;
        int f = 0, oy = 0, ox = 0;
        if (ok) {
          const uint32_t Lu = (uint32_t)L;
          f = (int)(Lu / (uint32_t)p.G);
          const uint32_t rem = Lu - (uint32_t)f * (uint32_t)p.G;
          const int U = (int)(rem / (uint32_t)p.GW), V = (int)(rem - (uint32_t)U * (uint32_t)p.GW);
          ok = U >= 1 && U <= p.Ho && V >= 1 && V <= p.Wo;
          oy = U - 1;
          ox = V - 1;
        }
        okm[m] = ok; fm[m] = f; oym[m] = oy; oxm[m] = ox;
        if (p.out_stem && ok && (p.stem_sub == 1 || (((oy | ox) & 1) == 0))) {
          __nv_bfloat16* o = p.out_stem + (((size_t)f * p.sub_oh + oy / p.stem_sub) * p.sub_ow + ox / p.stem_sub) * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            reinterpret_cast<uint4*>(o)[q] = make_uint4(packed[m][4 * q], packed[m][4 * q + 1], packed[m][4 * q + 2], packed[m][4 * q + 3]);
        }
#pragma unroll
        for (int kc = 0; kc < 4; ++kc)
          *reinterpret_cast<uint4*>(sA2 + m * 8192 + (r >> 3) * 512 + kc * 128 + (r & 7) * 16) =
              make_uint4(packed[m][4 * kc], packed[m][4 * kc + 1], packed[m][4 * kc + 2], packed[m][4 * kc + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (r == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b = s2_smem_u32(sW1);
#pragma unroll
        for (int m = 0; m < S2_MT; ++m) {
          const uint32_t a = s2_smem_u32(sA2 + m * 8192);
          s2_umma(tmem_base + acc2_col + (uint32_t)(m * p.n1p), s2_desc(a, 128u, 512u), s2_desc(b, 128u, 512u), idesc1, 0u);
          s2_umma(tmem_base + acc2_col + (uint32_t)(m * p.n1p), s2_desc(a + 256, 128u, 512u), s2_desc(b + 256, 128u, 512u), idesc1, 1u);
        }
        s2_commit(c1_bar);
      }
      s2_wait(c1_bar, c1_phase);
      c1_phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int m = 0; m < S2_MT; ++m) {
        __nv_bfloat16* o = p.out_c1 + (((size_t)fm[m] * p.Ho + oym[m]) * p.Wo + oxm[m]) * p.n1;
        for (int c0 = 0; c0 < p.n1p; c0 += 16) {
          uint32_t v32[16];
          s2_ld16(tmem_lane + acc2_col + (uint32_t)(m * p.n1p + c0), v32);
          if (!okm[m]) continue;
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
      // (the next group's conv1 MMAs are issued after its bar.sync, i.e. after every thread's tcgen05.ld above)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace tdeed

extern "C" long long tdeed_stem_tc2_wimg_bytes(void) { return tdeed::S2_NPAIR * 1024; }

extern "C" int tdeed_stem_tc2_fwd(const void* frames_u8, int n_frames, int in_h, int in_w, int crop_y, int crop_x, int h, int w,
                                  int flip, const void* wimg, const float* b0, const float* pad_rgb_host, const void* w1_bf16,
                                  const float* b1, int n1, void* out_stem, int stem_sub, void* out_c1, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(frames_u8 && wimg && b0 && pad_rgb_host && w1_bf16 && b1 && out_c1, TDEED_ERR_SHAPE, "tdeed_stem_tc2_fwd: null pointer");
  TDEED_REQUIRE(n_frames > 0 && h > 0 && w > 0 && crop_y >= 0 && crop_x >= 0 && crop_y + h <= in_h && crop_x + w <= in_w, TDEED_ERR_SHAPE,
                "tdeed_stem_tc2_fwd: bad geometry n=%d in=%dx%d crop=(%d,%d) %dx%d", n_frames, in_h, in_w, crop_y, crop_x, h, w);
  TDEED_REQUIRE(n1 > 0 && n1 % 8 == 0 && n1 <= 64 && (stem_sub == 1 || stem_sub == 2), TDEED_ERR_SHAPE, "tdeed_stem_tc2_fwd: n1=%d stem_sub=%d", n1, stem_sub);
  Stem2Params p{};
  p.frames = (const uint8_t*)frames_u8;
  p.frames_bytes = (long long)n_frames * 3 * in_h * in_w;
  p.n = n_frames; p.in_h = in_h; p.in_w = in_w; p.crop_y = crop_y; p.crop_x = crop_x; p.H = h; p.W = w; p.flip = flip;
  p.Ho = (h + 1) / 2; p.Wo = (w + 1) / 2;
  p.GH = p.Ho + 1; p.GW = p.Wo + 1; p.G = p.GH * p.GW;
  p.total_pos = (long long)n_frames * p.G;
  TDEED_REQUIRE(p.total_pos < (1LL << 31) - S2_GROUP, TDEED_ERR_SHAPE, "tdeed_stem_tc2_fwd: too many positions");
  p.ngroups = (int)ceil_div_ll(p.total_pos, S2_GROUP);
  p.min_off = -p.GW - 1;
  p.npos = S2_GROUP + p.GW + 2;       // +1: the partner chunk of the unpaired centre tap reads one position further (zero weights)
  p.npos_pad = p.npos | 1;
  // tap pairs (dy,dx): plane = 2*(dy != 1) + (dx != 1); offset = -(dy == 0)*GW - (dx == 0)
  const int plane[S2_NPAIR] = {3, 3, 2, 1, 0};
  const int off[S2_NPAIR] = {-p.GW - 1, -1, -p.GW, -1, 0};
  const int lbo[S2_NPAIR] = {1, 1, p.GW, 1, 1};
  for (int q = 0; q < S2_NPAIR; ++q) { p.pair_plane[q] = plane[q]; p.pair_off[q] = off[q]; p.pair_lbo[q] = lbo[q]; }
  TDEED_REQUIRE(p.GW < 16384, TDEED_ERR_UNSUPPORTED, "tdeed_stem_tc2_fwd: frame too wide");
  p.wimg = (const uint8_t*)wimg; p.b0 = b0;
  auto bf = [](float v) { __nv_bfloat16 t = __float2bfloat16_rn(v); unsigned short u; memcpy(&u, &t, 2); return (uint32_t)u; };
  p.pad_rg = bf(pad_rgb_host[0]) | (bf(pad_rgb_host[1]) << 16);
  p.pad_b = bf(pad_rgb_host[2]);
  p.w1 = (const __nv_bfloat16*)w1_bf16; p.b1 = b1; p.n1 = n1; p.n1p = (n1 + 15) / 16 * 16;
  p.out_stem = (__nv_bfloat16*)out_stem; p.stem_sub = stem_sub;
  p.sub_oh = (p.Ho + stem_sub - 1) / stem_sub; p.sub_ow = (p.Wo + stem_sub - 1) / stem_sub;
  p.out_c1 = (__nv_bfloat16*)out_c1;
  p.tmem_cols = (64 + S2_MT * p.n1p) <= 128 ? 128 : 256;       // stem accumulators (2 x 32) + one conv1 accumulator per M tile
  const size_t smem = (size_t)S2_NPAIR * 1024 + 4096 + S2_MT * 8192 + S2_NBUF * (size_t)4 * p.npos_pad * 16 + (32 + 64) * sizeof(float) + 10 * sizeof(uint64_t) + 128;
  TDEED_REQUIRE(smem <= 227 * 1024, TDEED_ERR_UNSUPPORTED, "tdeed_stem_tc2_fwd: width %d needs %zu B of shared memory", w, smem);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_stem_tc2_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = 227 * 1024;
  }
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > (int)(512 / p.tmem_cols)) per_sm = (int)(512 / p.tmem_cols);
  if (per_sm < 1) per_sm = 1;
  int grid = kNumSMs * per_sm;
  if (grid > p.ngroups) grid = p.ngroups;
  stem_tc2_kernel<<<grid, S2_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("tdeed_stem_tc2_fwd");
}
