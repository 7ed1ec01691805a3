// (2b) tcgen05 GEMM with fused epilogue — the bf16 backend of tdeed_gemm_fwd.
//   out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] + residual[m, n] ),  A/W bf16, fp32 accumulate
//
// sm_100a structure: PERSISTENT CTAs (one per SM) loop over 128 x BLOCK_N output tiles; 6 warps, warp-specialised:
//   warp 0   TMA producer : cp.async.bulk.tensor loads of the A (128 x 64) and W (BLOCK_N x 64) k-blocks into a
//                           ring of up to 10 128B-swizzled shared-memory stages (mbarrier expect_tx /
//                           complete_tx).  The ring runs ahead across tile boundaries, which is what keeps
//                           enough bytes in flight for the thin-K (K = 24..56) layers that are pure HBM streams.
//   warp 1   MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16)
//                           four times per k-block into one of TWO TMEM accumulators; tcgen05.commit releases
//                           the smem stage / hands the accumulator to the epilogue.  Owns tcgen05.alloc/dealloc.
//   warps 2-5 epilogue    : tcgen05.ld 32x32b (lane = output row) -> bias / residual / activation in registers
//                           -> 16-byte global stores, overlapped with the next tile's loads and MMAs.
// A may be a virtual concat of up to two column segments (two tensor maps): the GatedShift concat; or a 4D
// strided view of an NHWC tensor (every `stride`-th pixel): the stride-2 1x1 shortcut conv as implicit GEMM.
// K tails and M / N tails rely on TMA out-of-bounds zero fill and masked stores.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstdio>

namespace tdeed {

constexpr int TC_BM = 128, TC_BK = 64, TC_MAX_STAGES = 10;
constexpr int TC_THREADS = 320;   // producer warp, MMA warp, 8 epilogue warps
constexpr int TC_THREADS16 = 576; // producer warp, MMA warp, 16 epilogue warps (EPI == 2)
constexpr int TC_THREADS_SC = TC_THREADS + 128;   // + 4 A-scaling warps (EPI == 3)

struct TcParams {
  long long M;
  int N, block_n, num_stages;
  int nkb[TDEED_GEMM_MAX_SEGS];      // k-blocks per segment
  int a_col0[TDEED_GEMM_MAX_SEGS];   // first column inside the segment's tensor
  int w_col0[TDEED_GEMM_MAX_SEGS];   // first column of W consumed by the segment
  int nseg;
  const float* bias;
  const void* residual;
  long long ldr;
  int res_dtype, act;
  void* out;
  long long ldo;
  int out_dtype;
  uint32_t tmem_cols;
  int m_tiles, n_tiles;
  // 4D gather geometry (strided 1x1 conv): an M tile is a bw x bh patch of output pixels of one frame
  int gather, Ho, Wo, bw, bh, tiles_x, tiles_y;
  int w_res;      // 1: the whole weight matrix (one n-tile, all k-blocks) is loaded once per CTA and stays in shared memory
  uint32_t w_kb_stride, w_res_bytes;
  int out_bufs;   // 1 or 2 output staging tiles
  int staged;     // 1: outputs go through the smem staging tile (coalesced stores); 0: row pieces straight from registers
  int fast;       // 1: specialised epilogue (bf16 out / bf16 residual, act none|relu): whole-row residual prefetch, pipelined TMEM loads
  int debug;   // TDEED_GEMM_DEBUG (dev only): 1 = skip global stores, 2 = skip TMEM loads, 4 = skip the MMAs
  // A-operand scaling (squeeze-excite fused into conv3, EPI == 3): A[m, k] *= a_scale[m / scale_rows][k] while the k-block sits
  // in shared memory, between the TMA load and the MMA
  const float* a_scale;
  int scale_rows, scale_ld, K;
  int wide_ld;   // 1: residual row pieces are 32-byte aligned: 256-bit loads
  int wide_st;   // 1: output row pieces are 32-byte aligned (ldo * 2 and the base multiples of 32): 256-bit stores
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (error code), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tdeed gemm_tc: mbarrier wait timed out (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                         // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;               // SBO = 1024 B   [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B   [61,64)
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


template <bool STAGED>
__device__ __forceinline__ void epi_fast_tile(const TcParams& p, uint64_t* full_bar, uint32_t parity, uint32_t tmem_row, int half,
                                              int ncols, int n0, long long m, bool row_ok, const float* s_bias, uint32_t srow_addr) {
  // the residual row pieces of the thread's first TWO chunks are requested before the accumulator wait (their latency hides
  // behind the MMA of this tile), the piece of chunk k+2 when chunk k has been consumed.  (One chunk at a time, requested
  // right before its use, exposed ~1 us of global latency per chunk: 24 % of the epilogue's stall samples.)
  uint4 rres[2][4];
  const bool has_res = p.residual != nullptr && row_ok;
  const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.ldr + n0;
  const bool wide_ld = p.wide_ld != 0;
  auto load_res = [&](uint4 (&dst)[4], int c0) {
#pragma unroll
    for (int q = 0; q < 4; q += 2) {
      dst[q] = dst[q + 1] = make_uint4(0u, 0u, 0u, 0u);
      if (wide_ld && has_res && c0 + 8 * q + 16 <= ncols) {
        ldg256(rp + c0 + 8 * q, dst[q], dst[q + 1]);
      } else {
        if (has_res && c0 + 8 * q < ncols) dst[q] = *reinterpret_cast<const uint4*>(rp + c0 + 8 * q);
        if (has_res && c0 + 8 * q + 8 < ncols) dst[q + 1] = *reinterpret_cast<const uint4*>(rp + c0 + 8 * q + 8);
      }
    }
  };
  load_res(rres[0], half * 32);
  load_res(rres[1], half * 32 + 64);
  mbar_wait(full_bar, parity);
  tcgen05_fence_after();
  const float lo = p.act == TDEED_ACT_RELU ? 0.f : -INFINITY;
  __nv_bfloat16* grow = (!STAGED && row_ok) ? reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + n0 : nullptr;
  const float* bias = s_bias + n0;
  uint32_t va[16], vb[16];
  // TMEM loads (16 columns each) run one piece ahead of the arithmetic; piece s covers columns c(s) = half*32 + 64*(s/2) + 16*(s%2)
  if (half * 32 < ncols) tmem_ld16(tmem_row + (uint32_t)(half * 32), va);
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int k = s >> 1, sub = s & 1;
    const int c0 = half * 32 + 64 * k + 16 * sub;
    const int cn = half * 32 + 64 * ((s + 1) >> 1) + 16 * ((s + 1) & 1);      // next piece
    if (c0 < ncols) {                          // uniform over the warp
      tmem_ld_wait();
      if (s & 1) {
        if (s < 7 && cn < ncols) tmem_ld16(tmem_row + (uint32_t)cn, va);
        epi_fast_chunk<STAGED>(vb, rres[k & 1][2], rres[k & 1][3], bias, c0, ncols, lo, srow_addr, grow, p.wide_st != 0);
        if (k < 2) load_res(rres[k & 1], c0 - 16 + 128);
      } else {
        if (cn < ncols) tmem_ld16(tmem_row + (uint32_t)cn, vb);
        epi_fast_chunk<STAGED>(va, rres[k & 1][0], rres[k & 1][1], bias, c0, ncols, lo, srow_addr, grow, p.wide_st != 0);
      }
    }
  }
}

// 16-warp variant: the thread's pieces are the 16-column pieces quarter*16 + 64k of its row.  One TMEM load in flight, the
// residual of the next piece requested while the current one is processed; latency is hidden by the other 15 warps.
__device__ __forceinline__ void epi_fast16_tile(const TcParams& p, uint64_t* full_bar, uint32_t parity, uint32_t tmem_row, int quarter,
                                                int ncols, int n0, long long m, bool row_ok, const float* s_bias) {
  const bool has_res = p.residual != nullptr && row_ok;
  const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.ldr + n0;
  uint4 r0[2], r1[2];
  const bool wide_ld = p.wide_ld != 0;
  auto load_res = [&](uint4 (&dst)[2], int c0) {
    dst[0] = dst[1] = make_uint4(0u, 0u, 0u, 0u);
    if (wide_ld && has_res && c0 + 16 <= ncols) {
      ldg256(rp + c0, dst[0], dst[1]);
    } else {
      if (has_res && c0 < ncols) dst[0] = *reinterpret_cast<const uint4*>(rp + c0);
      if (has_res && c0 + 8 < ncols) dst[1] = *reinterpret_cast<const uint4*>(rp + c0 + 8);
    }
  };
  load_res(r0, quarter * 16);
  mbar_wait(full_bar, parity);
  tcgen05_fence_after();
  const float lo = p.act == TDEED_ACT_RELU ? 0.f : -INFINITY;
  __nv_bfloat16* grow = row_ok ? reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + n0 : nullptr;
  const float* bias = s_bias + n0;
  uint32_t v[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c0 = quarter * 16 + 64 * k;
    if (c0 < ncols) {                          // uniform over the warp
      tmem_ld16(tmem_row + (uint32_t)c0, v);
      if (k & 1) {
        if (k < 3) load_res(r0, c0 + 64);
        tmem_ld_wait();
        epi_fast_chunk<false>(v, r1[0], r1[1], bias, c0, ncols, lo, 0u, grow, p.wide_st != 0);
      } else {
        if (k < 3) load_res(r1, c0 + 64);
        tmem_ld_wait();
        epi_fast_chunk<false>(v, r0[0], r0[1], bias, c0, ncols, lo, 0u, grow, p.wide_st != 0);
      }
    }
  }
}

// ---------------------------------------------------------------- kernel
struct TileCoord { int mt, nt; };

// EPI 0: generic epilogue (fp32 outputs, GELU, ...).  EPI 1: the specialised bf16 epilogue as a separate instantiation, so that the
// generic one keeps its own register allocation.
// EPI 2: the same straight-line epilogue on SIXTEEN warps (four per TMEM lane group, each taking every fourth 16-column piece;
// direct stores only).  The 8-warp epilogue issues one instruction per ~14 cycles and warp (a latency chain TMEM -> bias ->
// residual -> pack -> store with two warps per scheduler); twice the warps hide twice the latency.
// EPI 3: EPI 1 plus four "scaler" warps between the TMA producer and the MMA issuer: they multiply every A k-block by the
// per-(frame, channel) squeeze-excite gate in shared memory (one row per thread, fp32 multiply, round-to-nearest bf16 — the same
// arithmetic as the stand-alone se_scale pass, which cost one read + one write of the whole activation per block).
template <int EPI>
__global__ void __launch_bounds__(EPI == 2 ? TC_THREADS16 : (EPI == 3 ? TC_THREADS_SC : TC_THREADS), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle pattern
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_stage_bytes = TC_BM * TC_BK * 2;                 // 16 KB
  const uint32_t w_stage_bytes = (uint32_t)p.block_n * TC_BK * 2;   // block_n * 128 B
  // W-resident mode (s3-sized layers: W <= 72 KB, many M tiles per CTA): W occupies the head of the buffer and is fetched
  // ONCE; the ring then carries only the 16 KB A blocks.  Re-fetching W per tile was 55 % of the TMA fill traffic of these
  // layers, and the kernel runs close to the ~24 B/clk/SM shared-memory fill rate of TMA.
  uint8_t* w_region = smem;
  uint8_t* ring = smem + (p.w_res ? p.w_res_bytes : 0u);
  const uint32_t stage_bytes = p.w_res ? a_stage_bytes : a_stage_bytes + ((w_stage_bytes + 1023u) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.num_stages * stage_bytes);
  constexpr int EPIK = (EPI == 3) ? 1 : EPI;            // epilogue flavour
  uint64_t* w_bar = bars + 2 * TC_MAX_STAGES + 5;
  uint64_t* scaled_bar = bars + 2 * TC_MAX_STAGES + 6;   // [TC_MAX_STAGES] A k-block scaled in place (128 scaler threads)
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * TC_MAX_STAGES;        // [2]
  uint64_t* tmem_empty_bar = bars + 2 * TC_MAX_STAGES + 2;   // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 4);
  float* s_bias = reinterpret_cast<float*>(bars + 3 * TC_MAX_STAGES + 6);   // [n_tiles * block_n], zero padded
  long long* s_rowm_base = reinterpret_cast<long long*>(s_bias + p.n_tiles * p.block_n);       // [2][128] global row of a tile row
  uint8_t* s_out_base = reinterpret_cast<uint8_t*>(s_rowm_base + 2 * TC_BM);                     // [out_bufs][128][block_n*esz + 16]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = p.nkb[0] + (p.nseg > 1 ? p.nkb[1] : 0);
  const int num_tiles = p.m_tiles * p.n_tiles;
  const uint32_t a_tx_bytes = p.gather ? (uint32_t)(p.bw * p.bh) * 128u : a_stage_bytes;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    if (p.nseg > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a1)) : "memory");
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&scaled_bar[s], 128);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], EPI == 2 ? 512 : 256);     // every epilogue thread arrives
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.n_tiles * p.block_n; i += (int)blockDim.x) s_bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer (one lane) =====
    if (lane == 0) {
      uint32_t it = 0;
      if (p.w_res) {
        mbar_expect_tx(w_bar, (uint32_t)total_kb * w_stage_bytes);
        int kbi = 0;
        for (int s = 0; s < p.nseg; ++s)
          for (int i = 0; i < p.nkb[s]; ++i, ++kbi)
            tma_load_2d(&map_w, w_bar, w_region + (size_t)kbi * p.w_kb_stride, p.w_col0[s] + i * TC_BK, (int)(blockIdx.x % p.n_tiles) * p.block_n);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
        const int n0 = nt * p.block_n;
        int c1 = mt * TC_BM, c2 = 0, c3 = 0;
        if (p.gather) {
          const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y;
          c3 = mt / (p.tiles_x * p.tiles_y);
          c1 = tx * p.bw;
          c2 = ty * p.bh;
        }
        for (int s = 0; s < p.nseg; ++s) {
          const CUtensorMap* map_a = (s == 0) ? &map_a0 : &map_a1;
          for (int i = 0; i < p.nkb[s]; ++i, ++it) {
            const int stage = it % p.num_stages;
            const uint32_t round = it / p.num_stages;
            mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);
            uint8_t* sa = ring + (size_t)stage * stage_bytes;
            uint8_t* sw = sa + a_stage_bytes;
            mbar_expect_tx(&full_bar[stage], p.w_res ? a_tx_bytes : a_tx_bytes + w_stage_bytes);
            if (p.gather) tma_load_4d(map_a, &full_bar[stage], sa, p.a_col0[s] + i * TC_BK, c1, c2, c3);
            else tma_load_2d(map_a, &full_bar[stage], sa, p.a_col0[s] + i * TC_BK, c1);
            if (!p.w_res) tma_load_2d(&map_w, &full_bar[stage], sw, p.w_col0[s] + i * TC_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N = block_n, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    uint32_t it = 0, j = 0;
    if (p.w_res && blockIdx.x < num_tiles) mbar_wait(w_bar, 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const uint32_t acc = j & 1u;
      mbar_wait(&tmem_empty_bar[acc], ((j >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.block_n;
      for (int kb = 0; kb < total_kb; ++kb, ++it) {
        const int stage = it % p.num_stages;
        const uint32_t round = it / p.num_stages;
        mbar_wait(EPI == 3 ? &scaled_bar[stage] : &full_bar[stage], round & 1u);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(ring + (size_t)stage * stage_bytes);
          const uint32_t sw = p.w_res ? smem_u32(w_region) + (uint32_t)kb * p.w_kb_stride : sa + a_stage_bytes;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            if (p.debug & 4) break;
            const uint64_t adesc = umma_desc_sw128(sa + k * 32);
            const uint64_t bdesc = umma_desc_sw128(sw + k * 32);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                   // frees the smem stage once the MMAs retire
          if (kb == total_kb - 1) umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else if (EPI == 3 && warp >= 10) {
    // ===== A scaler (4 warps, thread = tile row): SE gate applied to the k-block in shared memory =====
    const int r = threadIdx.x - 320;                         // 0..127
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);                   // 128B swizzle: 16-byte chunk c of row r lives at chunk c ^ (r % 8)
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles;
      long long m = (long long)mt * TC_BM + r;
      if (m >= p.M) m = p.M - 1;                             // rows past the end are TMA zero fill: any finite gate does
      const float* gate = p.a_scale + (size_t)(m / p.scale_rows) * p.scale_ld;
      for (int kb = 0; kb < total_kb; ++kb, ++it) {
        const int stage = it % p.num_stages;
        const uint32_t round = it / p.num_stages;
        const int k0 = kb * TC_BK;
        // gate values first (L1 / L2 hits: the 128 rows of a tile belong to at most a handful of frames)
        float4 g[16];
#pragma unroll
        for (int q = 0; q < 16; ++q)
          g[q] = (k0 + 4 * q < p.K) ? __ldg(reinterpret_cast<const float4*>(gate + k0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(&full_bar[stage], round & 1u);
        const uint32_t base = smem_u32(ring + (size_t)stage * stage_bytes) + row_off;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (k0 + 8 * c >= p.K) break;                      // K tail: zero-filled chunks stay zero
          const uint32_t addr = base + (((uint32_t)c ^ sw) << 4);
          uint4 v = lds128(addr);
          const float4 ga = g[2 * c], gb = g[2 * c + 1];
          uint32_t w4[4] = {v.x, v.y, v.z, v.w};
          const float gs[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            w4[q] = pack_bf16x2(__uint_as_float(w4[q] << 16) * gs[2 * q], __uint_as_float(w4[q] & 0xffff0000u) * gs[2 * q + 1]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w4[0]), "r"(w4[1]), "r"(w4[2]), "r"(w4[3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async-proxy reads
        mbar_arrive(&scaled_bar[stage]);
      }
    }
  } else {
    // ===== epilogue: 8 warps (256 threads).  Warp w may only touch TMEM lanes [32*(w%4), +32); the two warps of a
    // lane group split the accumulator's 32-column chunks (even / odd).  A thread owns a ROW of the accumulator:
    //   phase 1  residual piece of the row (global, requested before the accumulator is ready) ; TMEM -> registers ;
    //            + bias (smem) + residual, activation -> shared-memory staging tile (row-private, conflict-free)
    //   phase 2  staging tile -> global with fully coalesced 16-byte stores.  (Writing row pieces straight to global
    //            costs one 32-byte sector per lane per instruction: measured at 25-40 % of the kernel time in r1c.)
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lg * 32 + lane;                           // row inside the tile
    const int et = threadIdx.x - 64;                        // 0..255
    const int esz = (p.out_dtype == TDEED_F32) ? 4 : 2;
    const int pitch = p.block_n * esz + 16;                 // (pitch/16) odd -> conflict-free 16-byte row accesses
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const uint32_t acc = j & 1u;
      const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
      const int n0 = nt * p.block_n;
      long long* s_rowm = s_rowm_base + acc * TC_BM;        // double buffered: the previous tile's phase 2 may still read its table
      uint8_t* s_out = s_out_base + (p.out_bufs > 1 ? (size_t)acc * TC_BM * pitch : 0);
      long long m;
      bool row_ok;
      if (p.gather) {
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, f = mt / (p.tiles_x * p.tiles_y);
        const int by = r / p.bw, bx = r - by * p.bw;
        const int oy = ty * p.bh + by, ox = tx * p.bw + bx;
        row_ok = (by < p.bh) && (oy < p.Ho) && (ox < p.Wo);
        m = ((long long)f * p.Ho + oy) * p.Wo + ox;
      } else {
        m = (long long)mt * TC_BM + r;
        row_ok = m < p.M;
      }
      const int ncols = min(p.block_n, p.N - n0);           // multiple of 8
      const uint32_t tmem_row = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * (uint32_t)p.block_n;
      if constexpr (EPI == 2) {
        epi_fast16_tile(p, &tmem_full_bar[acc], (j >> 1) & 1u, tmem_row, half /* = column quarter */, ncols, n0, m, row_ok, s_bias);
        tcgen05_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
        continue;
      }
      if (half == 0) s_rowm[r] = row_ok ? m : -1;           // global row of every tile row (or -1), for phase 2
      uint8_t* srow = s_out + r * pitch;
      if constexpr (EPIK == 1) {
        if (p.staged && p.out_bufs == 1 && j > 0) epi_bar_sync();   // single staging tile: previous phase 2 must have drained
        if (p.staged) epi_fast_tile<true>(p, &tmem_full_bar[acc], (j >> 1) & 1u, tmem_row, half, ncols, n0, m, row_ok, s_bias, smem_u32(srow));
        else epi_fast_tile<false>(p, &tmem_full_bar[acc], (j >> 1) & 1u, tmem_row, half, ncols, n0, m, row_ok, s_bias, 0u);
        tcgen05_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
        if (!p.staged) continue;
        epi_bar_sync();
        // phase 2: a warp copies 16 rows; lanes cover the 16-byte chunks of one row (or of several rows when rows are short)
        const int cpr = ncols >> 3;                           // <= 32
        const int rpp = 32 / cpr;
        const int sub = lane / cpr, ch = lane - sub * cpr;
        const int ew = warp - 2;
        if (sub < rpp) {
          const uint32_t s_out_addr = smem_u32(s_out);
          for (int rr = sub; rr < 16; rr += rpp) {
            const int row = ew * 16 + rr;
            const long long mm = s_rowm[row];
            if (mm >= 0)
              *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + ((size_t)mm * p.ldo + n0) * 2 + (size_t)ch * 16) =
                  lds128(s_out_addr + (uint32_t)(row * pitch + ch * 16));
          }
        }
        continue;
      } else {
      bool waited = false;
      if (p.staged && p.out_bufs == 1 && j > 0) epi_bar_sync();   // single staging tile: previous phase 2 must have drained
      for (int c0 = half * 32; c0 < ncols; c0 += 64) {
        // residual piece (up to 32 columns) requested first: its latency overlaps the barrier wait / TMEM load
        uint4 rraw[8];
        const bool has_res = p.residual != nullptr && row_ok;
        if (has_res) {
          if (esz == 4) {
            const float* rp = reinterpret_cast<const float*>(p.residual) + m * p.ldr + n0 + c0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + 4 * q < ncols) rraw[q] = *reinterpret_cast<const uint4*>(rp + 4 * q);
          } else {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.ldr + n0 + c0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (c0 + 8 * q < ncols) rraw[q] = *reinterpret_cast<const uint4*>(rp + 8 * q);
          }
        }
        if (!waited) {
          mbar_wait(&tmem_full_bar[acc], (j >> 1) & 1u);
          tcgen05_fence_after();
          waited = true;
        }
        uint32_t v32[32];
        if (!(p.debug & 2)) {
          tmem_ld32(tmem_row + (uint32_t)c0, v32);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v32[q] = 0u;
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int cl = c0 + 8 * h;                         // column inside the tile
          if (cl >= ncols) continue;
          float v[8];
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + n0 + cl);
          const float4 b1 = *reinterpret_cast<const float4*>(s_bias + n0 + cl + 4);
          v[0] = __uint_as_float(v32[8 * h + 0]) + b0.x; v[1] = __uint_as_float(v32[8 * h + 1]) + b0.y;
          v[2] = __uint_as_float(v32[8 * h + 2]) + b0.z; v[3] = __uint_as_float(v32[8 * h + 3]) + b0.w;
          v[4] = __uint_as_float(v32[8 * h + 4]) + b1.x; v[5] = __uint_as_float(v32[8 * h + 5]) + b1.y;
          v[6] = __uint_as_float(v32[8 * h + 6]) + b1.z; v[7] = __uint_as_float(v32[8 * h + 7]) + b1.w;
          if (has_res) {
            if (esz == 4) {
              const uint4 ra = rraw[2 * h], rb = rraw[2 * h + 1];
              v[0] += __uint_as_float(ra.x); v[1] += __uint_as_float(ra.y); v[2] += __uint_as_float(ra.z); v[3] += __uint_as_float(ra.w);
              v[4] += __uint_as_float(rb.x); v[5] += __uint_as_float(rb.y); v[6] += __uint_as_float(rb.z); v[7] += __uint_as_float(rb.w);
            } else {
              const uint4 ra = rraw[h];
              const uint32_t w4[4] = {ra.x, ra.y, ra.z, ra.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[2 * q] += __uint_as_float(w4[q] << 16);
                v[2 * q + 1] += __uint_as_float(w4[q] & 0xffff0000u);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = apply_act_rt(v[q], p.act);
          if (p.staged) {
            if (esz == 4) store8(reinterpret_cast<float*>(srow) + cl, v);
            else store8(reinterpret_cast<__nv_bfloat16*>(srow) + cl, v);
          } else if (row_ok && !(p.debug & 1)) {
            if (esz == 4) store8(reinterpret_cast<float*>(p.out) + m * p.ldo + n0 + cl, v);
            else store8(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + n0 + cl, v);
          }
        }
      }
      if (!waited) mbar_wait(&tmem_full_bar[acc], (j >> 1) & 1u);   // this warp had no chunk: still consume the phase
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);                    // accumulator may be overwritten by the MMA warp
      if (!p.staged) continue;                              // direct-store mode: rows were written from registers
      epi_bar_sync();                                       // staging tile + row table complete
      if (!(p.debug & 1)) {
        const int cpr = ncols * esz / 16;                   // 16-byte chunks per row
        const uint32_t cpr_magic = ((1u << 24) + (uint32_t)cpr - 1u) / (uint32_t)cpr;
        for (int idx = et; idx < TC_BM * cpr; idx += 256) {
          const int row = (int)(((unsigned long long)idx * cpr_magic) >> 24);
          const int ch = idx - row * cpr;
          const long long mm = s_rowm[row];
          if (mm >= 0)
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + ((size_t)mm * p.ldo + n0) * esz + (size_t)ch * 16) =
                *reinterpret_cast<const uint4*>(s_out + row * pitch + ch * 16);
        }
      }
      }   // generic epilogue
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 tensor of rank `rank` (dims[0] innermost, contiguous); strides in bytes for dims 1..rank-1; 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode_fn();
  TDEED_REQUIRE(enc != nullptr, TDEED_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   // rows narrower than the 128-byte box (K = 24..56 layers) are contiguous in memory: promoting every
                   // row request to 256 B made TMA pull 2.5x the tensor over the crossbar (ncu r1c) -> no promotion there
                   dims[0] * 2 >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                      : (dims[0] * 2 >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TDEED_REQUIRE(r == CUDA_SUCCESS, TDEED_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1]);
  return TDEED_OK;
}

static int make_map_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  return make_map(map, base, 2, dims, strides, box);
}

int gemm_tc_launch(long long M, int N, int K, int nseg, const tdeed_gemm_seg* segs, int gstride, int gh, int gw,
                   const void* W, const float* bias, const void* residual, long long ldr, int res_dtype, int act,
                   void* out, long long ldo, int out_dtype, cudaStream_t st, const float* a_scale, int scale_rows) {
  TDEED_REQUIRE(M > 0 && M < (1LL << 31) - TC_BM, TDEED_ERR_SHAPE, "gemm_tc: M=%lld out of range", M);
  TDEED_REQUIRE(N % 8 == 0 && K % 8 == 0, TDEED_ERR_SHAPE, "gemm_tc: N=%d, K=%d must be multiples of 8", N, K);
  TDEED_REQUIRE(!residual || res_dtype == out_dtype, TDEED_ERR_UNSUPPORTED,
                "gemm_tc: the residual must have the output's dtype (it shares the output staging tile)");
  TcParams p{};
  p.M = M; p.N = N; p.nseg = nseg;
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.res_dtype = res_dtype; p.act = act;
  p.out = out; p.ldo = ldo; p.out_dtype = out_dtype;
  p.a_scale = a_scale; p.scale_rows = scale_rows; p.scale_ld = K; p.K = K;
  p.wide_ld = (residual && res_dtype == TDEED_BF16 && (ldr * 2) % 32 == 0 && (reinterpret_cast<uintptr_t>(residual) & 31) == 0) ? 1 : 0;
  p.wide_st = (out_dtype == TDEED_BF16 && (ldo * 2) % 32 == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) ? 1 : 0;
  if (a_scale) {
    TDEED_REQUIRE(scale_rows > 0 && nseg == 1 && segs[0].col0 == 0 && gstride <= 1 && out_dtype == TDEED_BF16 &&
                  (act == TDEED_ACT_NONE || act == TDEED_ACT_RELU) && (reinterpret_cast<uintptr_t>(a_scale) & 15) == 0,
                  TDEED_ERR_UNSUPPORTED,
                  "gemm_tc: the A-scale prologue needs one un-gathered segment starting at column 0, bf16 output, act none|relu");
  }

  // M tiling
  long long frames = 0;
  if (gstride > 1) {
    p.gather = 1;
    p.Ho = (gh + gstride - 1) / gstride;
    p.Wo = (gw + gstride - 1) / gstride;
    TDEED_REQUIRE(M % ((long long)p.Ho * p.Wo) == 0, TDEED_ERR_SHAPE, "gemm_tc: M=%lld is not frames*%d*%d", M, p.Ho, p.Wo);
    frames = M / ((long long)p.Ho * p.Wo);
    p.bw = p.Wo < TC_BM ? p.Wo : TC_BM;
    p.bh = TC_BM / p.bw < p.Ho ? TC_BM / p.bw : p.Ho;
    p.tiles_x = ceil_div(p.Wo, p.bw);
    p.tiles_y = ceil_div(p.Ho, p.bh);
    p.m_tiles = (int)(frames * p.tiles_x * p.tiles_y);
  } else {
    p.m_tiles = (int)ceil_div_ll(M, TC_BM);
  }

  // tile width: one tile when N <= 256, else an even split; shrink for skinny-M problems to get more CTAs
  const int max_bn = (out_dtype == TDEED_F32) ? 128 : 256;   // fp32 staging tile: 128 x 128 x 4 B
  int n_tiles = ceil_div(N, max_bn);
  int block_n = ceil_div(ceil_div(N, n_tiles), 16) * 16;
  if (block_n < 32) block_n = 32;
  while (block_n > 32 && (long long)p.m_tiles * ceil_div(N, block_n) < kNumSMs) {
    const int nb = ceil_div(block_n / 2, 16) * 16;
    if (nb == block_n) break;
    block_n = nb;
  }
  // (r2 negative result: narrower n-tiles for the N = K = 368 layers — a 98 KB W slice and a 6-stage instead of a 3-stage A
  // ring — made them SLOWER, 160 -> 180 us and 281 -> 387 us with the A-scale prologue: the third pass over A from L2 and the
  // extra tiles cost more than the deeper ring buys.)
  p.n_tiles = ceil_div(N, block_n);
  p.block_n = block_n;
  uint32_t cols = 32;
  while ((int)cols < 2 * block_n + 16) cols <<= 1;     // two accumulators (+ slack: the epilogue reads 32-column chunks)
  if (cols > 512) cols = 512;
  p.tmem_cols = cols;

  CUtensorMap maps[3];
  int total_kb = 0, wcol = 0;
  for (int s = 0; s < nseg; ++s) {
    TDEED_REQUIRE(segs[s].lda % 8 == 0 && (reinterpret_cast<uintptr_t>(segs[s].a) & 15) == 0, TDEED_ERR_SHAPE,
                  "gemm_tc: segment %d needs lda %% 8 == 0 and a 16-byte aligned base", s);
    // TMA box origins must be 16-byte aligned in global memory: column offsets in A and in W are multiples of 8
    TDEED_REQUIRE(segs[s].col0 % 8 == 0 && wcol % 8 == 0, TDEED_ERR_SHAPE,
                  "gemm_tc: segment %d starts at A column %d / W column %d; both must be multiples of 8 (pad the segment)",
                  s, segs[s].col0, wcol);
    // the tensor spans columns [0, col0 + k): loads past it are zero-filled, which implements the K tail
    const long long cols_a = (long long)segs[s].col0 + segs[s].k;
    int rc;
    if (p.gather) {
      const long long C = segs[s].lda;
      cuuint64_t dims[4] = {(cuuint64_t)cols_a, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)frames};
      cuuint64_t strides[3] = {(cuuint64_t)gstride * C * 2, (cuuint64_t)gstride * gw * C * 2, (cuuint64_t)gh * gw * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1};
      rc = make_map(&maps[s], segs[s].a, 4, dims, strides, box);
    } else {
      rc = make_map_2d(&maps[s], segs[s].a, M, cols_a, segs[s].lda, TC_BM);
    }
    if (rc) return rc;
    p.nkb[s] = ceil_div(segs[s].k, TC_BK);
    p.a_col0[s] = segs[s].col0;
    p.w_col0[s] = wcol;
    wcol += segs[s].k;
    total_kb += p.nkb[s];
  }
  if (nseg == 1) maps[1] = maps[0];
  TDEED_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0, TDEED_ERR_SHAPE, "gemm_tc: W must be 16-byte aligned");
  int rc = make_map_2d(&maps[2], W, N, K, K, block_n);
  if (rc) return rc;

  const size_t w_kb_stride = ((size_t)block_n * TC_BK * 2 + 1023) & ~(size_t)1023;
  static int wres_env = -2;
  if (wres_env == -2) { const char* e = tdeed::dev_env("TDEED_GEMM_WRES"); wres_env = e ? atoi(e) : -1; }
  // one n-tile: W <= 72 KB stays resident next to a deep ring.  Several n-tiles (s4-sized layers, W ~ 280 KB): every CTA keeps the
  // SLICE of its own n-tile (tile index % n_tiles is constant per CTA when the grid is a multiple of n_tiles) — up to 150 KB,
  // with a 3-stage A ring and direct (unstaged) stores; A is then read n_tiles times but W no longer once per M tile.
  const size_t w_slice = (size_t)total_kb * w_kb_stride;
  const bool big_slice = w_slice > 72 * 1024;
  p.w_res = (w_slice <= 150 * 1024 && (p.n_tiles == 1 || kNumSMs / p.n_tiles >= 1) && p.m_tiles >= 4 * kNumSMs) ? 1 : 0;
  if (wres_env >= 0) p.w_res = p.w_res && wres_env;
  p.w_kb_stride = (uint32_t)w_kb_stride;
  p.w_res_bytes = p.w_res ? (uint32_t)(total_kb * w_kb_stride) : 0u;
  const size_t stage_bytes = p.w_res ? (size_t)TC_BM * TC_BK * 2 : (size_t)TC_BM * TC_BK * 2 + w_kb_stride;
  int stages = (int)((200 * 1024 - p.w_res_bytes) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) stages = 2;
  p.num_stages = stages;
  const size_t bias_bytes = (size_t)p.n_tiles * block_n * sizeof(float);
  // Staging pays one CTA-wide barrier per tile: measured (tools/gemm_micro.py) to win for wide rows without a
  // residual (s3/s4 conv1: -10..-25 %) and to lose for thin rows or when the residual read already pulled the
  // row's lines into L1.
  static int dbg = -1;
  if (dbg < 0) { const char* e = tdeed::dev_env("TDEED_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  static int fast_env = -1;
  if (fast_env < 0) { const char* e = tdeed::dev_env("TDEED_GEMM_FAST_EPI"); fast_env = e ? atoi(e) : 1; }
  p.fast = (fast_env && dbg == 0 && out_dtype == TDEED_BF16 && (act == TDEED_ACT_NONE || act == TDEED_ACT_RELU) && block_n <= 256) ? 1 : 0;
  // With the specialised epilogue the direct row-piece stores win for one-n-tile and short-K layers (ncu per-launch times,
  // 57-clip batch: N = 152 conv1 198 vs 281 us, K 152 -> N 368 454 vs 571 us, strided shortcut convs 252 vs 370 us): no
  // CTA-wide barrier, no second pass over the tile.  The wide long-K layers without a residual (s4 conv1, K = N = 368) still
  // prefer the staged, coalesced stores (171 vs 245 us).
  const char* force_staged = tdeed::dev_env("TDEED_GEMM_STAGED");
  const bool direct_wins = p.fast && (N <= 256 || K <= 192);
  p.staged = force_staged ? atoi(force_staged) : (!direct_wins && residual == nullptr && N >= 96 ? 1 : 0);
  if (p.w_res && big_slice) p.staged = 0;
  const size_t stage_out_bytes = p.staged ? (size_t)TC_BM * ((size_t)block_n * (out_dtype == TDEED_F32 ? 4 : 2) + 16) : 0;
  const size_t fixed = 1024 + p.w_res_bytes + (3 * TC_MAX_STAGES + 6) * sizeof(uint64_t) + bias_bytes + 2 * TC_BM * sizeof(long long) + stage_out_bytes;
  TDEED_REQUIRE(fixed + 2 * stage_bytes <= 227 * 1024, TDEED_ERR_UNSUPPORTED, "gemm_tc: N=%d too wide for the bias staging area", N);
  // a second output staging tile saves one CTA-wide barrier per tile; take it when >= 3 ring stages still fit
  p.out_bufs = (fixed + stage_out_bytes + 3 * stage_bytes <= 227 * 1024) ? 2 : 1;
  const size_t fixed2 = fixed + (p.out_bufs > 1 ? stage_out_bytes : 0);
  while (stages > 2 && fixed2 + stages * stage_bytes > 227 * 1024) --stages;
  p.num_stages = stages;
  const size_t smem = fixed2 + stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int num_tiles = p.m_tiles * p.n_tiles;
  int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  if (p.w_res) grid -= grid % p.n_tiles;             // every CTA then sees one fixed n-tile: its resident W slice
  static int trace = -1;
  if (trace < 0) { const char* e = tdeed::dev_env("TDEED_GEMM_TRACE"); trace = e ? atoi(e) : 0; }
  if (trace)
    fprintf(stderr, "gemm_tc M=%lld N=%d K=%d nseg=%d gather=%d res=%d act=%d | block_n=%d n_tiles=%d w_res=%d stages=%d staged=%d out_bufs=%d fast=%d grid=%d smem=%zu\n",
            M, N, K, nseg, p.gather, residual != nullptr, act, block_n, p.n_tiles, p.w_res, stages, p.staged, p.out_bufs, p.fast, grid, smem);
  static int epi16_env = -1;
  if (epi16_env < 0) { const char* e = tdeed::dev_env("TDEED_GEMM_EPI16"); epi16_env = e ? atoi(e) : 0; }
  if (a_scale) {
    TDEED_REQUIRE(p.fast, TDEED_ERR_UNSUPPORTED, "gemm_tc: the A-scale prologue needs the specialised bf16 epilogue (N tile <= 256)");
    gemm_tc_kernel<3><<<grid, TC_THREADS_SC, smem, st>>>(maps[0], maps[1], maps[2], p);
    return check_launch("tdeed_gemm_scaled_fwd(tcgen05)");
  }
  if (p.fast && !p.staged && epi16_env) gemm_tc_kernel<2><<<grid, TC_THREADS16, smem, st>>>(maps[0], maps[1], maps[2], p);
  else if (p.fast) gemm_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], p);
  else gemm_tc_kernel<0><<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], p);
  return check_launch("tdeed_gemm_fwd(tcgen05)");
}

}  // namespace tdeed
