// (2b) tcgen05 GEMM with fused epilogue — the bf16 backend of tdeed_gemm_fwd.
//   out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] + residual[m, n] ),  A/W bf16, fp32 accumulate
//
// sm_100a structure (one output tile of 128 x BLOCK_N per CTA, 6 warps, warp-specialised):
//   warp 0   TMA producer : cp.async.bulk.tensor 2D loads of the A (128 x 64) and W (BLOCK_N x 64)
//                           k-blocks into a ring of 128B-swizzled shared-memory stages (mbarrier
//                           expect_tx / complete_tx)
//   warp 1   MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128,
//                           N=BLOCK_N, K=16) four times per k-block, accumulator in TMEM;
//                           tcgen05.commit releases the smem stage / signals the epilogue.
//                           This warp also owns tcgen05.alloc / dealloc.
//   warps 2-5 epilogue    : tcgen05.ld 32x32b (lane = output row) -> bias / residual / activation in
//                           registers -> 16-byte global stores.
// A may be a virtual concat of up to two column segments (two tensor maps): the GatedShift concat.
// K tails and M / N tails rely on TMA out-of-bounds zero fill and masked stores.
#include "common.cuh"
#include <cuda.h>

namespace tdeed {

constexpr int TC_BM = 128, TC_BK = 64, TC_MAX_STAGES = 6;
constexpr int TC_THREADS = 192;

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(TC_THREADS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle pattern
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_stage_bytes = TC_BM * TC_BK * 2;                 // 16 KB
  const uint32_t w_stage_bytes = (uint32_t)p.block_n * TC_BK * 2;   // block_n * 128 B (block_n % 16 == 0 -> 1024-aligned for %8)
  const uint32_t stage_bytes = a_stage_bytes + ((w_stage_bytes + 1023u) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * TC_MAX_STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * p.block_n;
  const int total_kb = p.nkb[0] + (p.nseg > 1 ? p.nkb[1] : 0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    if (p.nseg > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a1)) : "memory");
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int kb = 0;
      for (int s = 0; s < p.nseg; ++s) {
        const CUtensorMap* map_a = (s == 0) ? &map_a0 : &map_a1;
        for (int i = 0; i < p.nkb[s]; ++i, ++kb) {
          const int stage = kb % p.num_stages;
          const uint32_t round = (uint32_t)(kb / p.num_stages);
          mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sw = sa + a_stage_bytes;
          mbar_expect_tx(&full_bar[stage], a_stage_bytes + w_stage_bytes);
          tma_load_2d(map_a, &full_bar[stage], sa, p.a_col0[s] + i * TC_BK, m0);
          tma_load_2d(&map_w, &full_bar[stage], sw, p.w_col0[s] + i * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N = block_n, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int kb = 0; kb < total_kb; ++kb) {
      const int stage = kb % p.num_stages;
      const uint32_t round = (uint32_t)(kb / p.num_stages);
      mbar_wait(&full_bar[stage], round & 1u);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sw = sa + a_stage_bytes;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint64_t adesc = umma_desc_sw128(sa + k * 32);
          const uint64_t bdesc = umma_desc_sw128(sw + k * 32);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);                   // frees the smem stage once the MMAs retire
        if (kb == total_kb - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    const int lg = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const long long m = (long long)m0 + lg * 32 + lane;
    const bool row_ok = m < p.M;
    for (int c0 = 0; c0 < p.block_n; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + c0 + 8 * h;
        if (n >= p.N) continue;          // N is a multiple of 8
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 * h + j]);
        if (p.bias) {
          const float4 b0 = *reinterpret_cast<const float4*>(p.bias + n);
          const float4 b1 = *reinterpret_cast<const float4*>(p.bias + n + 4);
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (p.residual) {
          float rv[8];
          if (p.res_dtype == TDEED_F32) load8(reinterpret_cast<const float*>(p.residual) + m * p.ldr + n, rv);
          else load8(reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.ldr + n, rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += rv[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = apply_act_rt(v[j], p.act);
        if (p.out_dtype == TDEED_F32) store8(reinterpret_cast<float*>(p.out) + m * p.ldo + n, v);
        else store8(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldo + n, v);
      }
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

// 2D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols], 128B swizzle.
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  TDEED_REQUIRE(enc != nullptr, TDEED_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TDEED_REQUIRE(r == CUDA_SUCCESS, TDEED_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return TDEED_OK;
}

int gemm_tc_launch(long long M, int N, int K, int nseg, const tdeed_gemm_seg* segs, const void* W, const float* bias,
                   const void* residual, long long ldr, int res_dtype, int act, void* out, long long ldo,
                   int out_dtype, cudaStream_t st) {
  TDEED_REQUIRE(M > 0 && M < (1LL << 31) - TC_BM, TDEED_ERR_SHAPE, "gemm_tc: M=%lld out of range", M);
  TDEED_REQUIRE(N % 8 == 0 && K % 8 == 0, TDEED_ERR_SHAPE, "gemm_tc: N=%d, K=%d must be multiples of 8", N, K);
  TcParams p{};
  p.M = M; p.N = N; p.nseg = nseg;
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.res_dtype = res_dtype; p.act = act;
  p.out = out; p.ldo = ldo; p.out_dtype = out_dtype;

  // tile width: one tile when N <= 256, else an even split; shrink for skinny-M problems to get more CTAs
  const int m_tiles = (int)ceil_div_ll(M, TC_BM);
  int n_tiles = ceil_div(N, 256);
  int block_n = ceil_div(ceil_div(N, n_tiles), 16) * 16;
  while (block_n > 32 && (long long)m_tiles * ceil_div(N, block_n) < kNumSMs) {
    const int nb = ceil_div(block_n / 2, 16) * 16;
    if (nb == block_n) break;
    block_n = nb;
  }
  n_tiles = ceil_div(N, block_n);
  p.block_n = block_n;
  uint32_t cols = 32;
  while ((int)cols < block_n) cols <<= 1;
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
    int rc = make_map(&maps[s], segs[s].a, M, (long long)segs[s].col0 + segs[s].k, segs[s].lda, TC_BM);
    if (rc) return rc;
    p.nkb[s] = ceil_div(segs[s].k, TC_BK);
    p.a_col0[s] = segs[s].col0;
    p.w_col0[s] = wcol;
    wcol += segs[s].k;
    total_kb += p.nkb[s];
  }
  if (nseg == 1) maps[1] = maps[0];
  TDEED_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0, TDEED_ERR_SHAPE, "gemm_tc: W must be 16-byte aligned");
  int rc = make_map(&maps[2], W, N, K, K, block_n);
  if (rc) return rc;

  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (((size_t)block_n * TC_BK * 2 + 1023) & ~(size_t)1023);
  int stages = total_kb < 4 ? total_kb : 4;
  if (stages < 1) stages = 1;
  p.num_stages = stages;
  const size_t smem = 1024 + stages * stage_bytes + (2 * TC_MAX_STAGES + 2) * sizeof(uint64_t);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    smem_set = 227 * 1024;
  }
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles);
  gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(maps[0], maps[1], maps[2], p);
  return check_launch("tdeed_gemm_fwd(tcgen05)");
}

}  // namespace tdeed
