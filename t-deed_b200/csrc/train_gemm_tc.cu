// tcgen05 weight-gradient GEMM:  out[i, j] = alpha * sum_r A[r, i] * B[r, j]   (A = dY [R, m], B = X [R, n], bf16)
//
// The reduction runs over the ROWS of two row-major activations, so both UMMA operands are "MN-major": the tile that
// TMA drops into shared memory for a box of 64 channels x 64 rows (128-byte swizzle) is exactly the canonical MN-major
// SWIZZLE_128B layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units — 64 contiguous channels per row, 8-row groups
// 1024 B apart (SBO), 64-channel blocks one box (8 KB) apart (LBO).  No transposes are materialised anywhere.
//   grid = (row splits, n tiles, m tiles); a CTA owns one 128 x n_tile output tile for its slice of rows:
//   warp 0   TMA producer (ring of stages, mbarrier expect_tx)
//   warp 1   MMA issuer: 4 x tcgen05.mma.kind::f16 (M=128, N=n_tile, K=16, a_major = b_major = MN) per 64-row stage
//   warps 2-5 epilogue: tcgen05.ld 32x32b -> fp32 partial tile in the workspace; a second kernel sums the row splits in
//            a fixed order (deterministic).
// Rows past R and channels past m / n are zero-filled by TMA.
#include "common.cuh"
#include <cuda.h>

namespace tdeed {

constexpr int WT_THREADS = 192;
constexpr int WT_KR = 64;                 // rows (K) per stage
constexpr int WT_BOX_BYTES = WT_KR * 128; // one 64-channel x 64-row box
constexpr int WT_BM = 128;

struct WtParams {
  long long R, rows_per_split;
  int m, n, n_tile, n_boxes, stages;
  uint32_t tmem_cols;
  float* part;            // [splits][m][n]
};

__device__ __forceinline__ uint32_t wt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wt_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void wt_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool wt_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(wt_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wt_mbar_wait(uint64_t* bar, uint32_t parity) {   // bounded: a bug traps instead of hanging the GPU
  if (wt_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!wt_mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tdeed gemm_tn_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void wt_tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(wt_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(wt_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// MN-major, 128-byte-swizzled operand: LBO = distance between 64-element MN blocks, SBO = distance between 8-row K groups
__device__ __forceinline__ uint64_t wt_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void wt_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void wt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(wt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wt_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(WT_THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const WtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = (2 + p.n_boxes) * WT_BOX_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* acc_bar = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, nt = blockIdx.y, mt = blockIdx.z;
  const long long r0 = (long long)split * p.rows_per_split;
  const long long r1 = min(p.R, r0 + p.rows_per_split);
  const int iters = (int)((r1 - r0 + WT_KR - 1) / WT_KR);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      wt_mbar_init(&full[s], 1);
      wt_mbar_init(&empty[s], 1);
    }
    wt_mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wt_smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        if (it >= p.stages) wt_mbar_wait(&empty[s], ((it / p.stages) - 1) & 1);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        wt_mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
        const int row = (int)(r0 + (long long)it * WT_KR);
        for (int b = 0; b < 2; ++b) wt_tma_load_2d(&map_a, &full[s], st + b * WT_BOX_BYTES, mt * WT_BM + b * 64, row);
        for (int b = 0; b < p.n_boxes; ++b) wt_tma_load_2d(&map_b, &full[s], st + (2 + b) * WT_BOX_BYTES, nt * p.n_tile + b * 64, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_tile >> 3) << 17) |
                             ((uint32_t)(WT_BM >> 4) << 24);
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        wt_mbar_wait(&full[s], (it / p.stages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = wt_smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b0 = a0 + 2 * WT_BOX_BYTES;
#pragma unroll
        for (int k = 0; k < WT_KR / 16; ++k) {
          const uint64_t ad = wt_desc_mn_sw128(a0 + k * 2048, WT_BOX_BYTES, 1024);
          const uint64_t bd = wt_desc_mn_sw128(b0 + k * 2048, WT_BOX_BYTES, 1024);
          wt_umma(tmem_base, ad, bd, idesc, (it | k) != 0 ? 1u : 0u);
        }
        wt_commit(&empty[s]);
      }
      wt_commit(acc_bar);
    }
  } else {
    wt_mbar_wait(acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lg = warp & 3;                         // TMEM lane group this warp may access
    const int i = mt * WT_BM + lg * 32 + lane;       // output row (A channel)
    float* orow = p.part + ((size_t)split * p.m + i) * p.n + (size_t)nt * p.n_tile;
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
      uint32_t v[16];
      wt_tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (i < p.m) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = nt * p.n_tile + c0 + j;
          if (col < p.n) orow[c0 + j] = __uint_as_float(v[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

static __global__ void wt_split_reduce_kernel(const float* __restrict__ part, int splits, long long split_stride, int m, int n,
                                              float* __restrict__ out, long long ldo, float alpha) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)m * n) return;
  const int i = (int)(idx / n), j = (int)(idx - (long long)i * n);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(size_t)z * split_stride + idx];
  out[(size_t)i * ldo + j] = alpha * s;
}

typedef CUresult (*WtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static WtEncodeFn wt_encode_fn() {
  static WtEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WtEncodeFn>(ptr);
  }
  return fn;
}

static int wt_make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld) {
  WtEncodeFn enc = wt_encode_fn();
  TDEED_REQUIRE(enc != nullptr, TDEED_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)WT_KR};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   cols * 2 >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (cols * 2 >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TDEED_REQUIRE(r == CUDA_SUCCESS, TDEED_ERR_CUDA, "gemm_tn_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return TDEED_OK;
}

struct WtPlan {
  int m_tiles, n_tiles, n_tile, n_boxes, stages, splits;
  long long rows_per_split;
  size_t smem;
  uint32_t tmem_cols;
};

static WtPlan wt_plan(long long R, int m, int n) {
  WtPlan pl;
  pl.m_tiles = ceil_div(m, WT_BM);
  pl.n_tiles = ceil_div(n, 256);
  pl.n_tile = ceil_div(ceil_div(n, pl.n_tiles), 16) * 16;
  pl.n_boxes = ceil_div(pl.n_tile, 64);
  const int stage_bytes = (2 + pl.n_boxes) * WT_BOX_BYTES;
  pl.stages = (200 * 1024) / stage_bytes;
  if (pl.stages > 8) pl.stages = 8;
  pl.smem = (size_t)pl.stages * stage_bytes + (2 * pl.stages + 1) * 8 + 16 + 1024;
  uint32_t cols = 32;
  while ((int)cols < pl.n_tile) cols <<= 1;
  pl.tmem_cols = cols;
  const int tiles = pl.m_tiles * pl.n_tiles;
  long long splits = ceil_div_ll(2 * kNumSMs, tiles);
  const long long by_rows = ceil_div_ll(R, 8 * WT_KR);           // at least 512 rows per split
  if (splits > by_rows) splits = by_rows;
  if (splits < 1) splits = 1;
  long long rps = ceil_div_ll(ceil_div_ll(R, splits), WT_KR) * WT_KR;
  pl.rows_per_split = rps;
  pl.splits = (int)ceil_div_ll(R, rps);
  return pl;
}

bool gemm_tn_tc_applicable(int a_dtype, int b_dtype, const void* A, long long lda, const void* B, long long ldb, long long R,
                           int gather_stride) {
  return a_dtype == TDEED_BF16 && b_dtype == TDEED_BF16 && gather_stride == 1 && lda % 8 == 0 && ldb % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 && R >= 4 * WT_KR &&
         R < (1LL << 31);
}

long long gemm_tn_tc_workspace_floats(long long R, int m, int n) {
  return (long long)wt_plan(R, m, n).splits * m * n;
}

int gemm_tn_tc_launch(const void* A, long long lda, const void* B, long long ldb, long long R, int m, int n, float alpha,
                      float* out, long long ldo, float* ws, cudaStream_t st) {
  const WtPlan pl = wt_plan(R, m, n);
  CUtensorMap map_a, map_b;
  int rc = wt_make_map(&map_a, A, R, m, lda);
  if (rc) return rc;
  rc = wt_make_map(&map_b, B, R, n, ldb);
  if (rc) return rc;
  WtParams p{};
  p.R = R; p.rows_per_split = pl.rows_per_split; p.m = m; p.n = n; p.n_tile = pl.n_tile; p.n_boxes = pl.n_boxes;
  p.stages = pl.stages; p.tmem_cols = pl.tmem_cols; p.part = ws;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "gemm_tn_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(pl.splits, pl.n_tiles, pl.m_tiles);
  gemm_tn_tc_kernel<<<grid, WT_THREADS, pl.smem, st>>>(map_a, map_b, p);
  rc = check_launch("tdeed_gemm_tn(tcgen05)");
  if (rc) return rc;
  wt_split_reduce_kernel<<<(unsigned)ceil_div_ll((long long)m * n, 256), 256, 0, st>>>(ws, pl.splits, (long long)m * n, m, n, out, ldo, alpha);
  return check_launch("tdeed_gemm_tn(tcgen05 reduce)");
}

}  // namespace tdeed
