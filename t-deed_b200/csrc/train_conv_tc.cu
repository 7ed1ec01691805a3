// tcgen05 weight gradient of the grouped 3x3 convolution (bf16 activations):
//   dW[co][ci][ky][kx] = sum_{n,oy,ox} dY[n,oy,ox,co] * X[n, s*oy+ky-1, s*ox+kx-1, g*gw+ci]
//
// Per tap this is a dY^T X product over the pixels — the MN-major tcgen05 GEMM of train_gemm_tc.cu — restricted to the
// block diagonal (co and ci of the same group).  A CTA takes a block of 128 channels and one kernel row ky and runs
// three accumulations (kx = 0,1,2), each a full 128 x 128 tile in TMEM (3 x 128 = 384 columns): the tensor core
// computes 8x (gw=16) / 16x (gw=8) more products than needed, which is still ~30x cheaper than the CUDA-core kernel it
// replaces because the layer is bound by operand delivery, not by MMA issue.  Operands come straight from the NHWC
// tensors through 4D TMA boxes of 64 channels x (bw x bh = 64 pixels): the three shifted X windows are three boxes of
// the same tensor map at x-1 / x / x+1 (out-of-bounds pixels — the conv padding — are zero-filled by TMA); stride 2 uses
// four parity views of X (even/odd rows x even/odd columns) so that every tap is again a dense box.
//   grid = (pixel-tile splits, 3 kernel rows, channel blocks);  warp 0 TMA producer, warp 1 MMA issuer, warps 2-5
//   epilogue (each lane = one output channel picks its group's 16 / 8 columns and writes them to the partial buffer);
//   partials over the splits are summed in a fixed order by a second kernel (deterministic).
#include "train_reduce.cuh"
#include <cuda.h>

namespace tdeed {

constexpr int CT_THREADS = 192;
constexpr int CT_KR = 64;
constexpr int CT_BOX_BYTES = CT_KR * 128;
constexpr int CT_STAGE_BYTES = 8 * CT_BOX_BYTES;    // dY: 2 boxes, X: 3 taps x 2 boxes

struct CtParams {
  int n, oh, ow, c, gw, stride;
  int bw, bh, tiles_x, tiles_y;
  long long num_tiles, tiles_per_split;
  int stages;
  float* part;      // [splits][c][gw][9]
};

__device__ __forceinline__ uint32_t ct_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ct_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ct_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool ct_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(ct_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void ct_mbar_wait(uint64_t* bar, uint32_t parity) {
  if (ct_mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!ct_mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tdeed conv3x3g_bwd_weight_tc: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void ct_tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ct_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(ct_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint64_t ct_desc_mn_sw128(uint32_t smem_addr) {   // LBO = one box (next 64 channels), SBO = 8 rows
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(CT_BOX_BYTES >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void ct_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void ct_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ct_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ct_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

struct CtMaps {
  CUtensorMap dy;
  CUtensorMap x[4];     // stride 1: x[0];  stride 2: parity views [py*2 + px]
};

__global__ void __launch_bounds__(CT_THREADS, 1)
conv3x3g_bwd_weight_tc_kernel(const __grid_constant__ CtMaps maps, const CtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * CT_STAGE_BYTES);
  uint64_t* empty = full + p.stages;
  uint64_t* acc_bar = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, ky = blockIdx.y, cb = blockIdx.z;
  const int c0 = cb * 128;
  const long long t0 = (long long)split * p.tiles_per_split;
  const long long t1 = min(p.num_tiles, t0 + p.tiles_per_split);
  const int iters = (int)(t1 - t0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ct_mbar_init(&full[s], 1);
      ct_mbar_init(&empty[s], 1);
    }
    ct_mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // y tap: stride 1 -> row offset ky-1 in the single view; stride 2 -> parity py and offset dy
      const int py = p.stride == 2 ? (ky == 1 ? 0 : 1) : 0;
      const int dy = p.stride == 2 ? (ky == 0 ? -1 : 0) : ky - 1;
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        if (it >= p.stages) ct_mbar_wait(&empty[s], ((it / p.stages) - 1) & 1);
        uint8_t* st = smem + (size_t)s * CT_STAGE_BYTES;
        ct_mbar_expect_tx(&full[s], (uint32_t)CT_STAGE_BYTES);
        const long long tile = t0 + it;
        const int tx = (int)(tile % p.tiles_x), ty = (int)((tile / p.tiles_x) % p.tiles_y);
        const int f = (int)(tile / ((long long)p.tiles_x * p.tiles_y));
        const int ox0 = tx * p.bw, oy0 = ty * p.bh;
        for (int b = 0; b < 2; ++b) ct_tma_load_4d(&maps.dy, &full[s], st + b * CT_BOX_BYTES, c0 + b * 64, ox0, oy0, f);
        for (int kx = 0; kx < 3; ++kx) {
          const int px = p.stride == 2 ? (kx == 1 ? 0 : 1) : 0;
          const int dx = p.stride == 2 ? (kx == 0 ? -1 : 0) : kx - 1;
          const CUtensorMap* mx = &maps.x[py * 2 + px];
          for (int b = 0; b < 2; ++b)
            ct_tma_load_4d(mx, &full[s], st + (2 + kx * 2 + b) * CT_BOX_BYTES, c0 + b * 64, ox0 + dx, oy0 + dy, f);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      for (int it = 0; it < iters; ++it) {
        const int s = it % p.stages;
        ct_mbar_wait(&full[s], (it / p.stages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = ct_smem_u32(smem + (size_t)s * CT_STAGE_BYTES);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const uint32_t b0 = a0 + (2 + kx * 2) * CT_BOX_BYTES;
#pragma unroll
          for (int k = 0; k < CT_KR / 16; ++k)
            ct_umma(tmem_base + kx * 128, ct_desc_mn_sw128(a0 + k * 2048), ct_desc_mn_sw128(b0 + k * 2048), idesc, (it | k) != 0 ? 1u : 0u);
        }
        ct_commit(&empty[s]);
      }
      ct_commit(acc_bar);
    }
  } else {
    ct_mbar_wait(acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lg = warp & 3;
    const int cl = lg * 32 + lane;                 // channel inside the block = TMEM lane
    const int co = c0 + cl;
    const int half = lane >> 4;                    // which 16-channel pair of this warp's 32 lanes
    const int in_pair = cl & 15;
    // columns of this lane's group inside its pair's 16: gw == 16 -> all 16; gw == 8 -> the 8 of its own group
    const int j0 = p.gw == 16 ? 0 : (in_pair & 8);
    float* orow = p.part + ((size_t)split * p.c + co) * p.gw * 9 + ky * 3;
    for (int kx = 0; kx < 3; ++kx) {
      uint32_t v0[16], v1[16];
      const uint32_t base = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(kx * 128 + lg * 32);
      ct_tmem_ld16(base, v0);            // columns of the pair of lanes 0-15
      ct_tmem_ld16(base + 16, v1);       // columns of the pair of lanes 16-31
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (co < p.c && iters > 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float val = __uint_as_float(half ? v1[j] : v0[j]);
          const int ci = j - j0;
          if (ci >= 0 && ci < p.gw) orow[ci * 9 + kx] = val;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*CtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CtEncodeFn ct_encode_fn() {
  static CtEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<CtEncodeFn>(ptr);
  }
  return fn;
}

// 4D view (c, w', h', n) of an NHWC bf16 tensor with pixel strides sx / sy (in elements), box 64 x bw x bh x 1
static int ct_make_map(CUtensorMap* map, const void* base, long long c, long long wv, long long hv, long long n, long long sx,
                       long long sy, long long sn, int bw, int bh) {
  CtEncodeFn enc = ct_encode_fn();
  TDEED_REQUIRE(enc != nullptr, TDEED_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)wv, (cuuint64_t)hv, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TDEED_REQUIRE(r == CUDA_SUCCESS, TDEED_ERR_CUDA, "conv3x3g_bwd_weight_tc: cuTensorMapEncodeTiled failed (%d) dims=%lld,%lld,%lld,%lld",
                (int)r, c, wv, hv, n);
  return TDEED_OK;
}

struct CtPlan {
  int bw, bh, tiles_x, tiles_y, cblocks, splits, stages;
  long long num_tiles, tiles_per_split;
  size_t smem;
};

static CtPlan ct_plan(int n, int oh, int ow, int c) {
  CtPlan pl;
  int bw = 1;
  while (bw < ow && bw < CT_KR) bw <<= 1;
  pl.bw = bw;
  pl.bh = CT_KR / bw;
  pl.tiles_x = ceil_div(ow, pl.bw);
  pl.tiles_y = ceil_div(oh, pl.bh);
  pl.num_tiles = (long long)n * pl.tiles_x * pl.tiles_y;
  pl.cblocks = ceil_div(c, 128);
  long long splits = ceil_div_ll(2 * kNumSMs, 3 * pl.cblocks);
  const long long by_tiles = ceil_div_ll(pl.num_tiles, 8);
  if (splits > by_tiles) splits = by_tiles;
  if (splits < 1) splits = 1;
  pl.tiles_per_split = ceil_div_ll(pl.num_tiles, splits);
  pl.splits = (int)ceil_div_ll(pl.num_tiles, pl.tiles_per_split);
  pl.stages = 3;
  pl.smem = (size_t)pl.stages * CT_STAGE_BYTES + (2 * pl.stages + 1) * 8 + 16 + 1024;
  return pl;
}

bool conv3x3g_bwd_weight_tc_applicable(int dtype, const void* x, const void* dy, int c) {
  return dtype == TDEED_BF16 && c % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
}

long long conv3x3g_bwd_weight_tc_workspace_floats(int n, int h, int w, int c, int gw, int stride) {
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;
  return (long long)ct_plan(n, oh, ow, c).splits * c * gw * 9;
}

int conv3x3g_bwd_weight_tc_launch(const void* x, const void* dy, int n, int h, int w, int c, int gw, int stride, float* dw, float* ws,
                                  cudaStream_t st) {
  const int oh = (h + stride - 1) / stride, ow = (w + stride - 1) / stride;
  const CtPlan pl = ct_plan(n, oh, ow, c);
  CtMaps maps;
  int rc = ct_make_map(&maps.dy, dy, c, ow, oh, n, c, (long long)ow * c, (long long)oh * ow * c, pl.bw, pl.bh);
  if (rc) return rc;
  const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
  if (stride == 1) {
    rc = ct_make_map(&maps.x[0], xb, c, w, h, n, c, (long long)w * c, (long long)h * w * c, pl.bw, pl.bh);
    if (rc) return rc;
    maps.x[1] = maps.x[2] = maps.x[3] = maps.x[0];
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const long long wv = (w - px + 1) / 2, hv = (h - py + 1) / 2;     // pixels px, px+2, ... < w
        if (wv <= 0 || hv <= 0) {       // degenerate (w == 1 or h == 1): the view is empty -> any valid map, all loads out of bounds
          maps.x[py * 2 + px] = maps.dy;
          continue;
        }
        rc = ct_make_map(&maps.x[py * 2 + px], xb + ((size_t)py * w + px) * c, c, wv, hv, n, 2LL * c, 2LL * w * c, (long long)h * w * c,
                         pl.bw, pl.bh);
        if (rc) return rc;
      }
  }
  CtParams p{};
  p.n = n; p.oh = oh; p.ow = ow; p.c = c; p.gw = gw; p.stride = stride;
  p.bw = pl.bw; p.bh = pl.bh; p.tiles_x = pl.tiles_x; p.tiles_y = pl.tiles_y;
  p.num_tiles = pl.num_tiles; p.tiles_per_split = pl.tiles_per_split; p.stages = pl.stages; p.part = ws;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3g_bwd_weight_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    TDEED_REQUIRE(e == cudaSuccess, TDEED_ERR_CUDA, "conv3x3g_bwd_weight_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(pl.splits, 3, pl.cblocks);
  conv3x3g_bwd_weight_tc_kernel<<<grid, CT_THREADS, pl.smem, st>>>(maps, p);
  rc = check_launch("tdeed_conv3x3g_bwd_weight(tcgen05)");
  if (rc) return rc;
  const long long count = (long long)c * gw * 9;
  launch_partial_sum(ws, pl.splits, count, dw, st);
  return check_launch("tdeed_conv3x3g_bwd_weight(tcgen05 final)");
}

}  // namespace tdeed
