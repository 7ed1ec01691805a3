// Straight-line bf16 epilogue pieces shared by the tcgen05 GEMM kernels (gemm_tc.cu, gemm_thin.cu).
#pragma once
#include "common.cuh"

namespace tdeed {

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// two floats -> bf16x2 with ReLU in the conversion: max(x, 0) then round-to-nearest, the same values as fmaxf + cvt
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// 256-bit global store / load (sm_100: STG.E.ENL2.256 / LDG.E.ENL2.256); the address must be 32-byte aligned
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}

// One 16-column accumulator piece of one row -> + bias (+ bf16 residual), clamp at `lo` (0 = ReLU, -inf = none), bf16, store.
// Straight-line code (no runtime dtype / activation switches), so the 8-column groups interleave in the schedule.
// wide (kernel-uniform): every row piece is 32-byte aligned (row pitch and column origin multiples of 32 B): the 16 columns
// leave as ONE 256-bit store — half the store instructions and L2 write requests of the two 16-byte row pieces.
template <bool STAGED, bool HAS_RES = true>
__device__ __forceinline__ void epi_fast_chunk(const uint32_t (&acc)[16], const uint4& r0, const uint4& r1, const float* __restrict__ bias,
                                               int c0, int ncols, float lo, uint32_t srow_addr, __nv_bfloat16* grow, bool wide = false) {
  uint4 o2[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int cl = c0 + 8 * h;
    o2[h] = make_uint4(0u, 0u, 0u, 0u);
    if (cl < ncols) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + cl);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + cl + 4);
      const uint4 rv = h ? r1 : r0;
      const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
      float v[8];
      v[0] = __uint_as_float(acc[8 * h + 0]) + b0.x; v[1] = __uint_as_float(acc[8 * h + 1]) + b0.y;
      v[2] = __uint_as_float(acc[8 * h + 2]) + b0.z; v[3] = __uint_as_float(acc[8 * h + 3]) + b0.w;
      v[4] = __uint_as_float(acc[8 * h + 4]) + b1.x; v[5] = __uint_as_float(acc[8 * h + 5]) + b1.y;
      v[6] = __uint_as_float(acc[8 * h + 6]) + b1.z; v[7] = __uint_as_float(acc[8 * h + 7]) + b1.w;
      if (HAS_RES) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {         // residual words are zero when there is no residual
          v[2 * q] += __uint_as_float(w4[q] << 16);
          v[2 * q + 1] += __uint_as_float(w4[q] & 0xffff0000u);
        }
      }
      uint4 o;
      if (lo == 0.f) {                        // ReLU rides on the conversion (F2FP.RELU): no separate max per element
        o.x = pack_bf16x2_relu(v[0], v[1]);
        o.y = pack_bf16x2_relu(v[2], v[3]);
        o.z = pack_bf16x2_relu(v[4], v[5]);
        o.w = pack_bf16x2_relu(v[6], v[7]);
      } else {
        o.x = pack_bf16x2(fmaxf(v[0], lo), fmaxf(v[1], lo));
        o.y = pack_bf16x2(fmaxf(v[2], lo), fmaxf(v[3], lo));
        o.z = pack_bf16x2(fmaxf(v[4], lo), fmaxf(v[5], lo));
        o.w = pack_bf16x2(fmaxf(v[6], lo), fmaxf(v[7], lo));
      }
      o2[h] = o;
      if (STAGED) sts128(srow_addr + (uint32_t)cl * 2u, o);
      else if (grow && !(wide && c0 + 16 <= ncols)) *reinterpret_cast<uint4*>(grow + cl) = o;
    }
  }
  if (!STAGED && wide && grow && c0 + 16 <= ncols) stg256(grow + c0, o2[0], o2[1]);
}

}  // namespace tdeed
