// (2c) thin-K tcgen05 GEMM: the bf16 backend of tdeed_gemm_fwd for K <= 64 and N <= 256 (the stage-1/2 1x1 convs
// and s3.b1.conv1 of RegNetY-200MF: K = 24 / 32 / 56 with M up to 1.6e7 rows — pure HBM streams).
//
// Why not the TMA kernel (gemm_tc.cu): rows of 48..112 bytes make one TMA request each; measured with all math and
// stores disabled the TMA ring alone needed 2.0 us per 128 x 24 tile (tools/gemm_micro.py, TDEED_GEMM_DEBUG=7).  Here
//   * 4 producer warps stream the A tile with 16-byte cp.async (LDGSTS) — a warp reads 512 contiguous bytes — straight
//     into the canonical K-major NO-swizzle UMMA layout [8-column chunk][row][16 B] (SBO = 128 B, LBO = plane pitch),
//     and signal the stage with cp.async.mbarrier.arrive.noinc (mbarrier count 128);
//   * the weights (<= 256 x 64 bf16) are staged once per persistent CTA and stay resident;
//   * one MMA lane issues ceil(K/16) tcgen05.mma (M=128, N=block_n, K=16) per tile into one of two TMEM accumulators;
//   * 8 epilogue warps drain TMEM (tcgen05.ld.x32) -> bias / residual / activation -> 16-byte stores.
// A 16-stage ring of 8..16 KB stages keeps > 64 KB of loads in flight per SM.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include <cstdlib>

namespace tdeed {

constexpr int TH_BM = 128;
constexpr int TH_PROD_WARPS = 4;
constexpr int TH_THREADS = 32 * (TH_PROD_WARPS + 1 + 8);     // producers, MMA, epilogue
constexpr int TH_MAX_STAGES = 16;

struct ThinParams {
  long long M;
  int N, K, block_n;
  int nseg;
  const __nv_bfloat16* a[TDEED_GEMM_MAX_SEGS];
  long long lda[TDEED_GEMM_MAX_SEGS];
  int col0[TDEED_GEMM_MAX_SEGS];
  int kc[TDEED_GEMM_MAX_SEGS];     // 16-byte chunks per segment
  int kc_total;                    // real chunks (K / 8)
  int planes;                      // chunk planes per stage = 2 * ceil(K / 16)
  int num_stages, m_tiles;
  uint32_t a_plane_bytes, w_plane_bytes, stage_bytes;
  const __nv_bfloat16* W;
  const float* bias;
  const void* residual;
  long long ldr;
  int act;
  void* out;
  long long ldo;
  int out_dtype;
  uint32_t tmem_cols;
  int nacc_log2;   // log2 of the number of TMEM accumulators (2..8): thin tiles are cheap, the MMA warp may run that far ahead
  int wide_st; // output rows 32-byte aligned: 256-bit stores
  int fast;    // specialised bf16 epilogue (act none|relu): pipelined TMEM loads, cross-tile residual prefetch
  int mma_pair; // staged tiles the MMA warp handles per proxy fence (1..4)
  int split;   // fast && N <= 32: the two epilogue warp halves take alternate tiles (otherwise half of them would idle)
};

__device__ __forceinline__ uint32_t th_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void th_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(th_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool th_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(th_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void th_wait(uint64_t* bar, uint32_t parity) {
  if (th_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!th_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tdeed gemm_thin: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void th_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(th_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t th_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;          // Blackwell descriptor version; layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void th_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void th_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(th_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void th_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void th_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void th_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(TH_THREADS, 1)
gemm_thin_kernel(const ThinParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;                                                     // [planes][block_n rows][16 B] (+16 B pitch pad)
  uint8_t* sA = sW + (size_t)p.planes * p.w_plane_bytes;                  // [stages][planes][128 rows][16 B] (+16 B pad)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)p.num_stages * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TH_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * TH_MAX_STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * TH_MAX_STAGES + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TH_MAX_STAGES + 16);
  float* s_bias = reinterpret_cast<float*>(bars + 2 * TH_MAX_STAGES + 18);  // [block_n]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup: resident weights, zero padding planes, bias, barriers, TMEM ----
  for (int i = threadIdx.x; i < p.planes * p.block_n; i += TH_THREADS) {
    const int c = i / p.block_n, n = i - c * p.block_n;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < p.N && c < p.kc_total) v = *reinterpret_cast<const uint4*>(p.W + (size_t)n * p.K + c * 8);
    *reinterpret_cast<uint4*>(sW + (size_t)c * p.w_plane_bytes + n * 16) = v;
  }
  {  // chunk planes beyond K/8 (K % 16 == 8) are never written by the producers: zero them once in every stage
    const int npad = p.planes - p.kc_total;
    for (int i = threadIdx.x; i < p.num_stages * npad * TH_BM; i += TH_THREADS) {
      const int row = i % TH_BM, c = (i / TH_BM) % npad, st = i / (TH_BM * npad);
      *reinterpret_cast<uint4*>(sA + (size_t)st * p.stage_bytes + (size_t)(p.kc_total + c) * p.a_plane_bytes + row * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  for (int i = threadIdx.x; i < p.block_n; i += TH_THREADS) s_bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      th_mbar_init(&full_bar[s], 32 * TH_PROD_WARPS);       // one async arrival per producer thread
      th_mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < (1 << p.nacc_log2); ++a) {
      th_mbar_init(&tmem_full_bar[a], 1);
      th_mbar_init(&tmem_empty_bar[a], p.split ? 128 : 256);   // arrivals per accumulator drain
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TH_PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(th_smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // weights / zero planes were written by the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < TH_PROD_WARPS) {
    // ===== cp.async producers: 128 threads, chunk q = t + 128*i (i < kc_total) of the tile's 128 x kc_total 16-byte chunks =====
    // The (row, chunk) of a thread's i-th copy does not depend on the tile, so the shared-memory offset, the global element
    // offset and the segment are computed ONCE; per tile a copy is then an add and a cp.async.  (Recomputing them per chunk cost
    // ~66 instructions per 16 bytes and made these four warps the bottleneck of the whole kernel: ncu r1h, 464 producer
    // instructions per warp and tile while MMA and epilogue warps sat in their barrier waits.)
    const int t = threadIdx.x;
    const uint32_t sA_u32 = th_smem_u32(sA);
    uint32_t soff[8], goff[8], rowi[8];
    uint32_t seg1 = 0u;                       // bit i: the i-th copy reads segment 1
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int q = t + 32 * TH_PROD_WARPS * i;
      const int row = q / p.kc_total, c = q - row * p.kc_total;
      rowi[i] = (uint32_t)row;
      soff[i] = (uint32_t)c * p.a_plane_bytes + (uint32_t)row * 16u;
      if (c < p.kc[0]) {
        goff[i] = (uint32_t)(row * p.lda[0] + p.col0[0] + c * 8);
      } else {
        goff[i] = (uint32_t)(row * p.lda[1] + p.col0[1] + (c - p.kc[0]) * 8);
        seg1 |= 1u << i;
      }
    }
    const int n_it = p.kc_total;              // 128 * kc_total chunks over 128 threads
    const long long lda0 = p.lda[0], lda1 = p.lda[1];
    const __nv_bfloat16* a0 = p.a[0];
    const __nv_bfloat16* a1 = p.a[1];
    const int num_stages = p.num_stages, m_tiles = p.m_tiles;
    const uint32_t stage_bytes = p.stage_bytes;
    const long long M = p.M;
    int stage = 0;
    uint32_t round = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      th_wait(&empty_bar[stage], (round & 1u) ^ 1u);
      const long long m0 = (long long)tile * TH_BM;
      const uint32_t sbase = sA_u32 + (uint32_t)stage * stage_bytes;
      const __nv_bfloat16* b0 = a0 + m0 * lda0;
      const __nv_bfloat16* b1 = a1 + m0 * lda1;
      if (m0 + TH_BM <= M) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < n_it) {
            const __nv_bfloat16* src = ((seg1 >> i) & 1u ? b1 : b0) + goff[i];
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + soff[i]), "l"(src) : "memory");
          }
      } else {                                 // last tile: rows past M are zero-filled (src-size 0), never dereferenced
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < n_it) {
            const bool valid = m0 + rowi[i] < M;
            const __nv_bfloat16* src = valid ? ((seg1 >> i) & 1u ? b1 : b0) + goff[i] : a0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + soff[i]), "l"(src), "r"(valid ? 16 : 0) : "memory");
          }
      }
      // the mbarrier receives this thread's arrival once all of its cp.async above have landed
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(th_smem_u32(&full_bar[stage])) : "memory");
      if (++stage == num_stages) { stage = 0; ++round; }
    }
  } else if (warp == TH_PROD_WARPS) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(TH_BM >> 4) << 24);
    const uint32_t sA_u32 = th_smem_u32(sA), sW_u32 = th_smem_u32(sW);
    uint32_t it = 0;
    int stage = 0;
    uint32_t round = 0;
    const int num_stages = p.num_stages, m_tiles = p.m_tiles, nacc_log2 = p.nacc_log2;
    // Two tiles per trip: the generic->async proxy fence and the tcgen05 fence are paid once per PAIR of staged tiles (per
    // tile they kept this single warp ~70 % busy and made it the pacing role of the thin layers: ncu r1h).
    const int pair_env = p.mma_pair;
    for (int tile = blockIdx.x; tile < m_tiles; ) {
      int ntl = 1;                                   // tiles of this trip: up to p.mma_pair, limited by what is left
      while (ntl < pair_env && tile + ntl * (int)gridDim.x < m_tiles) ++ntl;
      uint32_t accs[4];
      int stages[4];
      for (int u = 0; u < ntl; ++u) {
        accs[u] = (it + u) & ((1u << nacc_log2) - 1u);
        th_wait(&tmem_empty_bar[accs[u]], (((it + u) >> nacc_log2) & 1u) ^ 1u);
        th_wait(&full_bar[stage], round & 1u);
        stages[u] = stage;
        if (++stage == num_stages) { stage = 0; ++round; }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) data -> visible to the MMA (async proxy)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        for (int u = 0; u < ntl; ++u) {
          const uint32_t a0 = sA_u32 + (uint32_t)stages[u] * p.stage_bytes;
          for (int k = 0; k < p.planes / 2; ++k) {
            const uint64_t adesc = th_desc(a0 + (uint32_t)(2 * k) * p.a_plane_bytes, p.a_plane_bytes, 128u);
            const uint64_t bdesc = th_desc(sW_u32 + (uint32_t)(2 * k) * p.w_plane_bytes, p.w_plane_bytes, 128u);
            th_umma(tmem_base + accs[u] * (uint32_t)p.block_n, adesc, bdesc, idesc, k != 0 ? 1u : 0u);
          }
          th_commit(&empty_bar[stages[u]]);
          th_commit(&tmem_full_bar[accs[u]]);
        }
      }
      __syncwarp();
      it += ntl;
      tile += ntl * (int)gridDim.x;
    }
  } else {
    // ===== epilogue (8 warps): TMEM -> bias / residual / activation -> global =====
    const int ew = warp - (TH_PROD_WARPS + 1);
    const int lg = warp & 3;
    const int half = ew >> 2;
    const int r = lg * 32 + lane;
    const int esz = (p.out_dtype == TDEED_F32) ? 4 : 2;
    if (p.fast) {
      // Specialised bf16 epilogue.  The residual piece of a thread's FIRST chunk of the NEXT tile is requested before the current
      // tile is processed, so its global latency overlaps a whole tile of work (with one tile in flight per warp it was fully
      // exposed: the accumulators are ready long before the epilogue gets to them); TMEM loads run one 16-column piece ahead.
      const float lo = p.act == TDEED_ACT_RELU ? 0.f : -INFINITY;
      const uint32_t it_step = p.split ? 2u : 1u;
      const int cbase = p.split ? 0 : half * 32;
      const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(p.residual);
      uint4 rres[2][4];
      auto load_res = [&](uint4 (&dst)[4], long long m, int c0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dst[q] = make_uint4(0u, 0u, 0u, 0u);
          if (res != nullptr && m < p.M && c0 + 8 * q < p.N) dst[q] = *reinterpret_cast<const uint4*>(res + m * p.ldr + c0 + 8 * q);
        }
      };
      uint32_t it = p.split ? (uint32_t)half : 0u;
      long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
      if (tile < p.m_tiles) load_res(rres[0], tile * TH_BM + r, cbase);
      for (; tile < p.m_tiles; tile += (long long)it_step * gridDim.x, it += it_step) {
        const uint32_t acc = it & ((1u << p.nacc_log2) - 1u);
        const long long m = tile * TH_BM + r;
        const bool row_ok = m < p.M;
        const uint32_t tmem_row = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * (uint32_t)p.block_n;
        __nv_bfloat16* grow = row_ok ? reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo : nullptr;
        load_res(rres[1], m, cbase + 64);
        th_wait(&tmem_full_bar[acc], (it >> p.nacc_log2) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t va[16], vb[16];
        if (cbase < p.N) th_ld16_nowait(tmem_row + (uint32_t)cbase, va);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int k = s >> 1;
          const int c0 = cbase + 64 * k + 16 * (s & 1);
          const int cn = cbase + 64 * ((s + 1) >> 1) + 16 * ((s + 1) & 1);
          if (c0 < p.N) {
            th_ld_wait();
            if (s & 1) {
              if (s < 7 && cn < p.N) th_ld16_nowait(tmem_row + (uint32_t)cn, va);
              epi_fast_chunk<false>(vb, rres[k & 1][2], rres[k & 1][3], s_bias, c0, p.N, lo, 0u, grow, p.wide_st != 0);
              if (k >= 1 && k < 2) load_res(rres[k & 1], m, c0 - 16 + 128);       // chunk 3 (rres[1]); chunk 2 is loaded below
            } else {
              if (cn < p.N) th_ld16_nowait(tmem_row + (uint32_t)cn, vb);
              epi_fast_chunk<false>(va, rres[k & 1][0], rres[k & 1][1], s_bias, c0, p.N, lo, 0u, grow, p.wide_st != 0);
            }
          }
          if (s == 1 && cbase + 128 < p.N) load_res(rres[0], m, cbase + 128);    // chunk 2 of this tile
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        th_arrive(&tmem_empty_bar[acc]);
        // first chunk of the next tile (after the last use of rres[0] in this one)
        const long long tnext = tile + (long long)it_step * gridDim.x;
        if (tnext < p.m_tiles) load_res(rres[0], tnext * TH_BM + r, cbase);
      }
    } else {
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & ((1u << p.nacc_log2) - 1u);
      const long long m = (long long)tile * TH_BM + r;
      const bool row_ok = m < p.M;
      const uint32_t tmem_row = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * (uint32_t)p.block_n;
      bool waited = false;
      for (int c0 = half * 32; c0 < p.N; c0 += 64) {
        uint4 rraw[8];
        const bool has_res = p.residual != nullptr && row_ok;
        if (has_res) {
          if (esz == 4) {
            const float* rp = reinterpret_cast<const float*>(p.residual) + m * p.ldr + c0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + 4 * q < p.N) rraw[q] = *reinterpret_cast<const uint4*>(rp + 4 * q);
          } else {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.ldr + c0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (c0 + 8 * q < p.N) rraw[q] = *reinterpret_cast<const uint4*>(rp + 8 * q);
          }
        }
        if (!waited) {
          th_wait(&tmem_full_bar[acc], (it >> p.nacc_log2) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          waited = true;
        }
        uint32_t v32[32];
        th_ld32(tmem_row + (uint32_t)c0, v32);
        if (!row_ok) continue;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int n = c0 + 8 * h;
          if (n >= p.N) continue;
          float v[8];
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + n);
          const float4 b1 = *reinterpret_cast<const float4*>(s_bias + n + 4);
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
          if (esz == 4) store8(reinterpret_cast<float*>(p.out) + m * p.ldo + n, v);
          else store8(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + n, v);
        }
      }
      if (!waited) th_wait(&tmem_full_bar[acc], (it >> p.nacc_log2) & 1u);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      th_arrive(&tmem_empty_bar[acc]);
    }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TH_PROD_WARPS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// true when the thin kernel covers the problem (bf16, single k-block, one n-tile, no gather)
bool gemm_thin_applicable(int N, int K, int nseg, const tdeed_gemm_seg* segs, int gstride) {
  if (gstride > 1 || K > 64 || K % 8 != 0 || N > 256 || N % 8 != 0) return false;
  for (int s = 0; s < nseg; ++s)
    if (segs[s].k % 8 != 0 || segs[s].col0 % 8 != 0 || segs[s].lda % 8 != 0 || (reinterpret_cast<uintptr_t>(segs[s].a) & 15) != 0) return false;
  return true;
}

int gemm_thin_launch(long long M, int N, int K, int nseg, const tdeed_gemm_seg* segs, const void* W, const float* bias,
                     const void* residual, long long ldr, int res_dtype, int act, void* out, long long ldo, int out_dtype,
                     cudaStream_t st) {
  TDEED_REQUIRE(M > 0 && M < (1LL << 31) - TH_BM, TDEED_ERR_SHAPE, "gemm_thin: M=%lld out of range", M);
  TDEED_REQUIRE(!residual || res_dtype == out_dtype, TDEED_ERR_UNSUPPORTED, "gemm_thin: the residual must have the output's dtype");
  TDEED_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0, TDEED_ERR_SHAPE, "gemm_thin: W must be 16-byte aligned");
  ThinParams p{};
  p.M = M; p.N = N; p.K = K; p.nseg = nseg;
  for (int s = 0; s < TDEED_GEMM_MAX_SEGS; ++s) {
    const tdeed_gemm_seg& g = segs[s < nseg ? s : 0];
    p.a[s] = (const __nv_bfloat16*)g.a; p.lda[s] = g.lda; p.col0[s] = g.col0; p.kc[s] = (s < nseg) ? g.k / 8 : 0;
  }
  if (nseg == 1) p.kc[0] = K / 8;
  p.kc_total = K / 8;
  p.planes = 2 * ((K + 15) / 16);
  p.block_n = (N + 15) / 16 * 16;
  if (p.block_n < 32) p.block_n = 32;
  p.a_plane_bytes = TH_BM * 16 + 16;            // +16 B: consecutive chunk planes start 4 banks apart
  p.w_plane_bytes = (uint32_t)p.block_n * 16 + 16;
  p.stage_bytes = (uint32_t)p.planes * p.a_plane_bytes;
  p.m_tiles = (int)ceil_div_ll(M, TH_BM);
  p.W = (const __nv_bfloat16*)W; p.bias = bias; p.residual = residual; p.ldr = ldr; p.act = act;
  p.out = out; p.ldo = ldo; p.out_dtype = out_dtype;
  int nacc_log2 = 1;
  while (nacc_log2 < 3 && (2 << nacc_log2) * p.block_n + 16 <= 512) ++nacc_log2;
  p.nacc_log2 = nacc_log2;
  uint32_t cols = 32;
  while ((int)cols < (1 << nacc_log2) * p.block_n + 16) cols <<= 1;
  if (cols > 512) cols = 512;
  p.tmem_cols = cols;
  const size_t fixed = (size_t)p.planes * p.w_plane_bytes + (2 * TH_MAX_STAGES + 18) * sizeof(uint64_t) + (size_t)p.block_n * sizeof(float) + 128;
  int stages = (int)((200 * 1024 - fixed) / p.stage_bytes);
  if (stages > TH_MAX_STAGES) stages = TH_MAX_STAGES;
  TDEED_REQUIRE(stages >= 2, TDEED_ERR_UNSUPPORTED, "gemm_thin: N=%d K=%d does not fit", N, K);
  p.num_stages = stages;
  const size_t smem = fixed + (size_t)stages * p.stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "gemm_thin: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  static int fast_env = -1;
  if (fast_env < 0) { const char* e = tdeed::dev_env("TDEED_GEMM_FAST_EPI"); fast_env = e ? atoi(e) : 1; }
  p.fast = (fast_env && out_dtype == TDEED_BF16 && (act == TDEED_ACT_NONE || act == TDEED_ACT_RELU)) ? 1 : 0;
  p.wide_st = (out_dtype == TDEED_BF16 && (ldo * 2) % 32 == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) ? 1 : 0;
  p.split = (p.fast && N <= 32) ? 1 : 0;
  static int pair_env = -1;
  if (pair_env < 0) { const char* e = tdeed::dev_env("TDEED_THIN_MMA_PAIR"); pair_env = e ? atoi(e) : 2; }
  p.mma_pair = p.num_stages >= 8 ? (pair_env > 4 ? 4 : (pair_env < 1 ? 1 : pair_env)) : 1;     // tiles per MMA-warp trip
  if (p.mma_pair > (1 << p.nacc_log2)) p.mma_pair = 1 << p.nacc_log2;   // a trip must not wait for an accumulator it fills itself
  const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
  gemm_thin_kernel<<<grid, TH_THREADS, smem, st>>>(p);
  return check_launch("tdeed_gemm_fwd(tcgen05 thin-K)");
}

}  // namespace tdeed
