// (3b) grouped 3x3 convolution (+folded BN + ReLU) on tcgen05 — the bf16 backend of timm Bottleneck.conv2
// (group width 8 or 16, stride 1 or 2, pad 1).
//
// Implicit GEMM without im2col.  Output pixels are enumerated in a zero-padded "position space": every frame is
// a (Ho+pad) x (Wo+pad) grid, so a 3x3 tap is a CONSTANT offset in linear position for every pixel of every frame.
// A CTA tile is 128 consecutive positions (= the 128 TMEM lanes / GEMM rows).  The input window of the tile is
// staged once in shared memory as planes [8-channel chunk][position][16 B] — which is exactly the canonical
// K-major no-swizzle UMMA operand layout with SBO = 128 B and LBO = plane pitch — so the A operand of tap
// (dy, dx) is the SAME shared-memory data addressed through a descriptor whose start address is shifted by the
// tap offset: 9 tcgen05.mma (M=128, N=16, K=16) per 16-channel pair accumulate the whole 3x3 kernel in TMEM.
// A pair is two width-8 groups (block-diagonal 16x16 weight tile) or one width-16 group (dense tile).
// Stride 2 uses four parity planes (even/odd row x even/odd column) so taps stay unit-stride.
// Weights of the CTA's channel block stay resident in shared memory; CTAs are persistent over position tiles,
// several per SM so that staging, MMA and epilogue of different tiles overlap.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include <cstdlib>

namespace tdeed {

constexpr int C3T_THREADS = 32 * (4 + 1 + 4);   // 4 producer warps, MMA warp, 4 epilogue warps
constexpr int C3T_MAX_PAIRS = 8;     // 128 channels per CTA

struct C3TParams {
  const __nv_bfloat16* in;
  __nv_bfloat16* out;
  const float* bias;
  const uint8_t* wimg;       // [pairs_total][9][512 B] canonical UMMA B tiles
  int n, H, W, C, Ho, Wo, stride, relu;
  int out_H, out_W, out_sy, out_sx, out_oy, out_ox;   // output tensor geometry: pixel (U-1, V-1) -> (sy*(U-1)+oy, sx*(V-1)+ox)
  int GH, GW, G;             // padded position grid per frame
  long long total_pos;
  int ntiles;
  int nplanes;
  int tap_plane[9], tap_off[9];
  int min_off, npos, npos_pad;
  int nchunks_real, pairs_total, pairs_blk;
  uint32_t tmem_cols;
  int wide_st;               // output rows are 32-byte aligned (C * 2 % 32 == 0): 256-bit stores
};

__device__ __forceinline__ uint32_t c3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t c3_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;          // Blackwell descriptor version; layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void c3_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void c3_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void c3_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void c3_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool c3_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(c3_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void c3_wait(uint64_t* bar, uint32_t parity) {
  if (c3_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!c3_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tdeed conv3x3g_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void c3_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c3_smem_u32(bar)) : "memory");
}

// Warp roles: warps 0-3 producers (position table + cp.async staging of the next tile's window into one of two
// input buffers), warp 4 MMA issuer, warps 5-8 epilogue (TMEM lanes = rows), two TMEM accumulators: staging of tile
// i+1, the MMAs of tile i and the epilogue of tile i-1 overlap inside one CTA.
__global__ void __launch_bounds__(C3T_THREADS)
conv3x3g_tc_kernel(const C3TParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair0 = blockIdx.y * p.pairs_blk;
  const int np = min(p.pairs_blk, p.pairs_total - pair0);
  const int nch = 2 * np;                        // staged chunk planes (a trailing zero plane pads odd chunk counts)
  const int chunk0 = 2 * pair0;
  const int nch_real = min(nch, p.nchunks_real - chunk0);
  const size_t in_bytes = (size_t)p.nplanes * 2 * p.pairs_blk * p.npos_pad * 16;
  uint8_t* sW = smem;                                              // [pairs_blk][9][512]
  uint8_t* sIn = sW + (size_t)p.pairs_blk * 9 * 512;               // [2][nplanes][nch][npos_pad][16]
  float* s_bias = reinterpret_cast<float*>(sIn + 2 * in_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + p.pairs_blk * 16);
  uint64_t* full_bar = bars;          // [2] input buffer staged          (128 async arrivals)
  uint64_t* empty_bar = bars + 2;     // [2] input buffer consumed        (tcgen05.commit)
  uint64_t* tfull_bar = bars + 4;     // [2] accumulator ready            (tcgen05.commit)
  uint64_t* tempty_bar = bars + 6;    // [2] accumulator drained          (128 epilogue threads)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 8);
  int* s_tbl = reinterpret_cast<int*>(bars + 10);                  // [2][npos][nplanes]

  for (int i = tid; i < np * 9 * 32; i += C3T_THREADS)
    reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(p.wimg + (size_t)pair0 * 9 * 512)[i];
  for (int i = tid; i < np * 16; i += C3T_THREADS) {
    const int ch = pair0 * 16 + i;
    s_bias[i] = (p.bias && ch < p.C) ? p.bias[ch] : 0.f;
  }
  // chunk planes beyond the tensor's channels are never staged: zero them once in both buffers
  for (int i = tid; i < 2 * p.nplanes * (nch - nch_real) * p.npos_pad; i += C3T_THREADS) {
    const int s = i % p.npos_pad;
    const int c = (i / p.npos_pad) % (nch - nch_real);
    const int pl = (i / (p.npos_pad * (nch - nch_real))) % p.nplanes;
    const int buf = i / (p.npos_pad * (nch - nch_real) * p.nplanes);
    *reinterpret_cast<uint4*>(sIn + buf * in_bytes + ((size_t)(pl * nch + nch_real + c) * p.npos_pad + s) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c3_smem_u32(&full_bar[i])), "r"(128));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c3_smem_u32(&empty_bar[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c3_smem_u32(&tfull_bar[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c3_smem_u32(&tempty_bar[i])), "r"(128));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(c3_smem_u32(s_tmem)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  const uint32_t acc_cols = (uint32_t)p.pairs_blk * 16u;
  const uint32_t plane_bytes = (uint32_t)p.npos_pad * 16u;

  if (warp < 4) {
    // ===== producers (128 threads) =====
    // A thread owns the window entries (position s, parity plane pl) with s*nplanes + pl = tid + 128k and copies ALL channel
    // chunks of such an entry: the pixel address is computed once per entry (and advanced incrementally from entry to entry:
    // no divisions inside the tile), a chunk is then two adds and a cp.async.  The former per-chunk scheme (position table in
    // shared memory + two divisions per 16 bytes) needed ~65 producer instructions per chunk and kept MMA and epilogue warps
    // waiting (ncu r1h: 85 % of the epilogue's samples sat in the accumulator wait).
    const int nplanes = p.nplanes;                         // 1 (stride 1) or 4 (stride 2)
    const int pl = tid & (nplanes - 1);
    const int s0 = nplanes == 4 ? tid >> 2 : tid;
    const int ds = 128 / nplanes;                          // position step between a thread's entries
    const int dF = ds / p.G, r1 = ds - dF * p.G, dU = r1 / p.GW, dV = r1 - dU * p.GW;
    const int py = pl >> 1, px = pl & 1;
    const int n_ent = s0 < p.npos ? (p.npos - s0 + ds - 1) / ds : 0;
    const uint32_t dst0 = (uint32_t)(pl * nch) * plane_bytes + (uint32_t)s0 * 16u;     // ((pl*nch + c)*npos_pad + s) * 16
    const uint32_t dst_step = (uint32_t)ds * 16u;
    const __nv_bfloat16* in0 = p.in + (size_t)chunk0 * 8;
    const int GW = p.GW, GH = p.GH, H = p.H, W = p.W, stride = p.stride;
    const long long total_pos = p.total_pos;
    const size_t C = (size_t)p.C;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      c3_wait(&empty_bar[buf], ((it >> 1) & 1u) ^ 1u);
      long long L = (long long)tile * 128 + p.min_off + s0;        // may be negative (halo before the first frame)
      const long long Ls = L + p.G;                                  // shifted by one frame: non-negative
      int f = (int)(Ls / p.G) - 1;
      const int rem = (int)(Ls - (long long)(f + 1) * p.G);
      int U = rem / GW, V = rem - U * GW;
      uint32_t dst = c3_smem_u32(sIn + buf * in_bytes) + dst0;
      for (int k = 0; k < n_ent; ++k) {
        const int iy = stride * (U - 1) + py, ix = stride * (V - 1) + px;
        const bool valid = f >= 0 && L < total_pos && U >= 1 && V >= 1 && iy < H && ix < W;
        const __nv_bfloat16* src = valid ? in0 + ((size_t)(f * H + iy) * W + ix) * C : in0;
        const int sz = valid ? 16 : 0;                               // zero-fill for padding
        uint32_t d = dst;
        for (int c = 0; c < nch_real; ++c, d += plane_bytes, src += 8)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");   // .ca: halo rows and parity neighbours are re-read by the next tiles / planes
        dst += dst_step;
        L += ds;
        V += dV;
        if (V >= GW) { V -= GW; ++U; }
        U += dU;
        if (U >= GH) { U -= GH; ++f; }
        f += dF;
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(c3_smem_u32(&full_bar[buf])) : "memory");
    }
  } else if (warp == 4) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t b0 = c3_smem_u32(sW);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      c3_wait(&tempty_bar[buf], ((it >> 1) & 1u) ^ 1u);
      c3_wait(&full_bar[buf], (it >> 1) & 1u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a0 = c3_smem_u32(sIn + buf * in_bytes);
        for (int pp = 0; pp < np; ++pp) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t a = a0 + ((uint32_t)(p.tap_plane[t] * nch + 2 * pp) * (uint32_t)p.npos_pad + (uint32_t)(p.tap_off[t] - p.min_off)) * 16u;
            const uint32_t b = b0 + (uint32_t)(pp * 9 + t) * 512u;
            c3_umma(tmem_base + buf * acc_cols + (uint32_t)pp * 16u, c3_desc(a, plane_bytes, 128u), c3_desc(b, 128u, 256u), idesc, t != 0 ? 1u : 0u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c3_smem_u32(&empty_bar[buf])) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c3_smem_u32(&tfull_bar[buf])) : "memory");
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue (4 warps): row = position; bias + ReLU -> bf16 NHWC =====
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const float lo = p.relu ? 0.f : -INFINITY;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      const long long L = (long long)tile * 128 + r;
      bool ok = L < p.total_pos;
      size_t obase = 0;
      if (ok) {
        const int Li = (int)L;
        const int f = Li / p.G;
        const int rem = Li - f * p.G;
        const int U = rem / p.GW, V = rem - U * p.GW;
        const int oy = p.out_sy * (U - 1) + p.out_oy, ox = p.out_sx * (V - 1) + p.out_ox;
        ok = U >= 1 && U <= p.Ho && V >= 1 && V <= p.Wo && oy < p.out_H && ox < p.out_W;
        obase = (((size_t)f * p.out_H + oy) * p.out_W + ox) * p.C + (size_t)pair0 * 16;
      }
      c3_wait(&tfull_bar[buf], (it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_lane = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * acc_cols;
      // TMEM loads run one channel pair ahead of the arithmetic; bias as 16-byte shared-memory reads, straight-line
      // bias / ReLU / pack per 8 channels (after the producer rewrite these four warps were the busiest: ncu r1h)
      __nv_bfloat16* grow = ok ? p.out + obase : nullptr;
      const int ncols = min(np * 16, p.C - pair0 * 16);
      const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
      uint32_t va[16], vb[16];
      c3_ld16_nowait(tmem_lane, va);
      for (int pp = 0; pp < np; pp += 2) {
        c3_ld_wait();
        if (pp + 1 < np) c3_ld16_nowait(tmem_lane + (uint32_t)(pp + 1) * 16u, vb);
        epi_fast_chunk<false, false>(va, zero4, zero4, s_bias, pp * 16, ncols, lo, 0u, grow, p.wide_st != 0);
        if (pp + 1 < np) {
          c3_ld_wait();
          if (pp + 2 < np) c3_ld16_nowait(tmem_lane + (uint32_t)(pp + 2) * 16u, va);
          epi_fast_chunk<false, false>(vb, zero4, zero4, s_bias, (pp + 1) * 16, ncols, lo, 0u, grow, p.wide_st != 0);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      c3_arrive(&tempty_bar[buf]);
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

struct C3TOut { int H, W, sy, sx, oy, ox; };

static int conv3x3g_tc_run(const void* in, int n, int h, int w, int c, int stride, const void* wimg, const float* bias, int relu,
                           void* out, void* stream, const C3TOut* og = nullptr) {
  using namespace tdeed;
  TDEED_REQUIRE(in && wimg && out, TDEED_ERR_SHAPE, "tdeed_conv3x3g_tc_fwd: null pointer");
  TDEED_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && (stride == 1 || stride == 2), TDEED_ERR_SHAPE,
                "tdeed_conv3x3g_tc_fwd: bad shape n=%d %dx%dx%d stride=%d", n, h, w, c, stride);
  C3TParams p{};
  p.in = (const __nv_bfloat16*)in; p.out = (__nv_bfloat16*)out; p.bias = bias; p.wimg = (const uint8_t*)wimg;
  p.n = n; p.H = h; p.W = w; p.C = c; p.stride = stride; p.relu = relu;
  p.Ho = (h + stride - 1) / stride; p.Wo = (w + stride - 1) / stride;
  if (og) { p.out_H = og->H; p.out_W = og->W; p.out_sy = og->sy; p.out_sx = og->sx; p.out_oy = og->oy; p.out_ox = og->ox; }
  else { p.out_H = p.Ho; p.out_W = p.Wo; p.out_sy = p.out_sx = 1; p.out_oy = p.out_ox = 0; }
  if (stride == 1) {
    // Shared padding: ONE zero row between consecutive frames and ONE zero column between consecutive rows.  Positions are
    // linear (U * GW + V, frames back to back), so the right neighbour of a row's last pixel IS the next row's pad column and
    // the row below a frame's last row IS the next frame's pad row: a (H+1) x (W+1) grid per frame instead of (H+2) x (W+2)
    // — 64 instead of 81 positions for the 7x7 frames of stage 4 (49 of them real: 77 % instead of 60 % useful MMA rows,
    // epilogue rows and staged window entries), 225 instead of 256 at 14x14.
    p.GH = h + 1; p.GW = w + 1; p.nplanes = 1;
    for (int t = 0; t < 9; ++t) { p.tap_plane[t] = 0; p.tap_off[t] = (t / 3 - 1) * p.GW + (t % 3 - 1); }
    p.min_off = -p.GW - 1;
    p.npos = 128 + 2 * p.GW + 2;
  } else {
    p.GH = p.Ho + 1; p.GW = p.Wo + 1; p.nplanes = 4;
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3, dx = t % 3;
      p.tap_plane[t] = ((dy != 1) ? 2 : 0) + ((dx != 1) ? 1 : 0);
      p.tap_off[t] = -(dy == 0 ? p.GW : 0) - (dx == 0 ? 1 : 0);
    }
    p.min_off = -p.GW - 1;
    p.npos = 128 + p.GW + 1;
  }
  p.npos_pad = p.npos | 1;                      // odd plane pitch: conflict-free 16-byte staging stores
  p.G = p.GH * p.GW;
  p.total_pos = (long long)n * p.G;
  const long long nt = ceil_div_ll(p.total_pos, 128);
  TDEED_REQUIRE(nt < (1LL << 31), TDEED_ERR_SHAPE, "tdeed_conv3x3g_tc_fwd: too many tiles");
  p.ntiles = (int)nt;
  p.nchunks_real = c / 8;
  p.wide_st = ((c * 2) % 32 == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) ? 1 : 0;
  p.pairs_total = (c + 15) / 16;
  // channel pairs per CTA: at most 8 (128 channels), fewer when the staged window of a wide frame would not fit
  const size_t per_pair = (size_t)9 * 512 + 2 * (size_t)p.nplanes * 2 * p.npos_pad * 16 + 64;     // two input buffers
  int max_pairs = (int)((200 * 1024 - 2 * (size_t)p.npos * p.nplanes * sizeof(int)) / per_pair);
  if (max_pairs > C3T_MAX_PAIRS) max_pairs = C3T_MAX_PAIRS;
  // Wide layers run better as several narrow channel blocks (more CTAs per SM: the staging / MMA / epilogue stages of more tiles
  // overlap) than as 128-channel blocks.  Measured on B200 at 5700 frames (us, max pairs 8 / 6 / 4 / 3 / 2):
  //   stride 1   7x7x368 : 191 / 148 / 148 / 153 / 161     14x14x320: 690 / 612 / 537 / 580 / 577     7x7x768: 385 / 330 / 328 / 330 / 365
  //              14x14x152: 321 / 320 / 336 / 381 / 445    28x28x128: 227 / 243 / 242 / 293 / 258  (narrow layers: keep 8)
  //   stride 2  14x14x368: 438 /  -  / 354 / 312 / 336     28x28x152: 682 / - / - / 598 / 761       14x14x768: 977 / - / 766 / 702 / 731
  //              56x56x128: 485 / - / - / 514 / 508  (<= 8 pairs: keep 8)
  // The choice depends on the layer only (never on the frame count): a batch and its clips one by one run the same kernel.
  if (stride == 1 && p.pairs_total >= 16 && max_pairs > 4) max_pairs = 4;
  if (stride == 2 && p.pairs_total >= 10 && max_pairs > 3) max_pairs = 3;
  static int pairs_env = -1;
  if (pairs_env < 0) { const char* e = tdeed::dev_env("TDEED_C3_MAX_PAIRS"); pairs_env = e ? atoi(e) : 0; }
  if (pairs_env > 0 && max_pairs > pairs_env) max_pairs = pairs_env;
  TDEED_REQUIRE(max_pairs >= 1, TDEED_ERR_UNSUPPORTED, "tdeed_conv3x3g_tc_fwd: frame width %d too large for the staged window", w);
  const int nblk = ceil_div(p.pairs_total, max_pairs);
  p.pairs_blk = ceil_div(p.pairs_total, nblk);
  uint32_t cols = 32;
  while ((int)cols < 2 * p.pairs_blk * 16) cols <<= 1;          // two accumulators
  p.tmem_cols = cols;
  const size_t smem = (size_t)p.pairs_blk * 9 * 512 + 2 * (size_t)p.nplanes * 2 * p.pairs_blk * p.npos_pad * 16 +
                      (size_t)p.pairs_blk * 16 * sizeof(float) + 10 * sizeof(uint64_t) + 2 * (size_t)p.npos * p.nplanes * sizeof(int) + 64;
  TDEED_REQUIRE((long long)p.npos * p.nplanes * 2 * p.pairs_blk < 65536, TDEED_ERR_UNSUPPORTED, "tdeed_conv3x3g_tc_fwd: tile too large");
  TDEED_REQUIRE(smem <= 227 * 1024, TDEED_ERR_UNSUPPORTED, "tdeed_conv3x3g_tc_fwd: width %d needs %zu B of shared memory", w, smem);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3g_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "tdeed_conv3x3g_tc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = 227 * 1024;
  }
  int per_sm = (int)((220 * 1024) / smem);
  if (per_sm > 512 / (int)cols) per_sm = 512 / (int)cols;
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  int gx = (kNumSMs * per_sm) / nblk;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  conv3x3g_tc_kernel<<<dim3(gx, nblk), C3T_THREADS, smem, (cudaStream_t)stream>>>(p);
  return check_launch("tdeed_conv3x3g_tc_fwd");
}

extern "C" int tdeed_conv3x3g_tc_fwd(const void* in, int n, int h, int w, int c, int stride, const void* wimg,
                                     const float* bias, void* out, void* stream) {
  TDEED_REQUIRE(bias, TDEED_ERR_SHAPE, "tdeed_conv3x3g_tc_fwd: null pointer");
  return conv3x3g_tc_run(in, n, h, w, c, stride, wimg, bias, 1, out, stream);
}

// training: raw convolution (no bias, no ReLU).  With a weight image built with transpose_flip = 1 (stride 1 only) this is
// also the DATA GRADIENT of the stride-1 grouped conv: dx = conv(dy, W^T flipped).
extern "C" int tdeed_conv3x3g_tc_raw_fwd(const void* in, int n, int h, int w, int c, int stride, const void* wimg, void* out,
                                         void* stream) {
  return conv3x3g_tc_run(in, n, h, w, c, stride, wimg, nullptr, 0, out, stream);
}

namespace tdeed {
// fp32 weights [C][gw][3][3] -> bf16 UMMA B tiles [ceil(C/16)][9 taps][16 out][16 in], each tile in the canonical K-major
// no-swizzle layout [n/8][k/8][n%8][k%8]; transpose_flip: tile(tap)[ci][co] = W[co][ci][8 - tap] (data-gradient kernel)
// transpose_flip >= 2: parity kernels of the STRIDE-2 data gradient, (py, px) = ((mode-2) >> 1, (mode-2) & 1): the gradient at
// input pixels (2j+py, 2i+px) is a stride-1 conv over the dy grid with taps at offsets {0} (parity 0: k = 1) or {0, +1}
// (parity 1: k = 2, 0) per axis; unused taps are zero.
__global__ void conv3_weight_image_kernel(const float* __restrict__ w, int C, int gw, int transpose_flip, __nv_bfloat16* __restrict__ img,
                                          int total) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = idx & 15, nrow = (idx >> 4) & 15, tap = (idx >> 8) % 9, pair = idx / (256 * 9);
  // tile element (n = output channel 16*pair + nrow, k = input channel 16*pair + k)
  const int cn = pair * 16 + nrow, ck = pair * 16 + k;
  float v = 0.f;
  if (cn < C && ck < C && cn / gw == ck / gw) {
    if (!transpose_flip) v = w[((size_t)cn * gw + (ck % gw)) * 9 + tap];
    else if (transpose_flip == 1) v = w[((size_t)ck * gw + (cn % gw)) * 9 + (8 - tap)];
    else {
      const int py = (transpose_flip - 2) >> 1, px = (transpose_flip - 2) & 1;
      const int a = tap / 3 - 1, b = tap % 3 - 1;                     // offsets on the dy grid
      const int ky = py == 0 ? (a == 0 ? 1 : -1) : (a == 0 ? 2 : (a == 1 ? 0 : -1));
      const int kx = px == 0 ? (b == 0 ? 1 : -1) : (b == 0 ? 2 : (b == 1 ? 0 : -1));
      if (ky >= 0 && kx >= 0) v = w[((size_t)ck * gw + (cn % gw)) * 9 + ky * 3 + kx];
    }
  }
  const int off = ((pair * 9 + tap) * 256) + ((nrow >> 3) * 2 + (k >> 3)) * 64 + (nrow & 7) * 8 + (k & 7);
  img[off] = __float2bfloat16_rn(v);
}
}  // namespace tdeed

extern "C" long long tdeed_conv3_weight_image_elems(int c) { return (long long)((c + 15) / 16) * 9 * 256; }

extern "C" int tdeed_conv3_weight_image(const float* weight, int c, int group_width, int transpose_flip, void* wimg, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(weight && wimg && c > 0 && (group_width == 8 || group_width == 16) && c % group_width == 0 && transpose_flip >= 0 &&
                transpose_flip <= 5, TDEED_ERR_SHAPE, "tdeed_conv3_weight_image: bad arguments");
  const int total = ((c + 15) / 16) * 9 * 256;
  conv3_weight_image_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(weight, c, group_width, transpose_flip,
                                                                                   (__nv_bfloat16*)wimg, total);
  return check_launch("tdeed_conv3_weight_image");
}

// Data gradient of the STRIDE-2 grouped conv on tcgen05: four stride-1 convolutions over the dy grid (one per input-pixel
// parity, weight images built with tdeed_conv3_weight_image mode 2 + 2*py + px), each scattering into dx at (2j+py, 2i+px).
// wimgs: the four images back to back (tdeed_conv3_weight_image_elems(c) elements each).  dx: NHWC [n, h, w, c] bf16.
extern "C" int tdeed_conv3x3g_tc_bwd_data_s2(const void* dy, int n, int h, int w, int c, const void* wimgs, void* dx, void* stream) {
  TDEED_REQUIRE(dy && wimgs && dx && n > 0 && h > 0 && w > 0, TDEED_ERR_SHAPE, "tdeed_conv3x3g_tc_bwd_data_s2: bad arguments");
  const int oh = (h + 1) / 2, ow = (w + 1) / 2;
  const long long img_bytes = tdeed_conv3_weight_image_elems(c) * 2;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      C3TOut og{h, w, 2, 2, py, px};
      int rc = conv3x3g_tc_run(dy, n, oh, ow, c, 1, (const uint8_t*)wimgs + (py * 2 + px) * img_bytes, nullptr, 0, dx, stream, &og);
      if (rc) return rc;
    }
  return TDEED_OK;
}
