// Deterministic per-channel (column) reductions shared by the training kernels.
#pragma once
#include "common.cuh"

namespace tdeed {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_GRID = kNumSMs * 4;

struct BnLayout {
  int c8n, rpi;   // channel octets; rows handled per CTA iteration
};
__host__ __device__ inline BnLayout bn_layout(int C) {
  BnLayout l;
  l.c8n = (C + 7) / 8;
  l.rpi = BN_THREADS / l.c8n;
  return l;
}

// Final stage of the column reductions: one WARP per channel sums the per-CTA partials part[p][which][ch] (lane l takes
// p = l, l+32, ... serially, then a fixed xor-shuffle tree) in double.  Deterministic; ~20x faster than one thread per
// channel walking all partials.  Call with ch uniform across the warp.
__device__ inline void warp_partial_sums(const float* __restrict__ part, int nparts, int cpad, int ch, double& s, double& q) {
  const int lane = threadIdx.x & 31;
  double a = 0.0, b = 0.0;
#pragma unroll 4
  for (int p = lane; p < nparts; p += 32) {
    a += (double)part[((size_t)p * 2) * cpad + ch];
    b += (double)part[((size_t)p * 2 + 1) * cpad + ch];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  s = a;
  q = b;
}

// Generic column reduction: OP::row(v-arrays) -> two accumulators per channel.  part: [grid][2][c8n*8]
template <typename T, typename OP>
__global__ void __launch_bounds__(BN_THREADS) bn_reduce_kernel(OP op, long long M, int C, float* __restrict__ part) {
  __shared__ float s_acc[BN_THREADS * 16];
  const BnLayout l = bn_layout(C);
  const int c8 = threadIdx.x % l.c8n, lr = threadIdx.x / l.c8n;
  const bool active = lr < l.rpi;
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  if (active) {
    const int nch = min(8, C - c8 * 8);
    op.begin(c8 * 8, nch);
    const long long stride = (long long)gridDim.x * l.rpi;
    long long r = (long long)blockIdx.x * l.rpi + lr;
    if constexpr (OP::kBatch > 1) {
      // kBatch rows per trip: all their loads are issued before the first accumulate, so a thread keeps kBatch x 16-32 B in
      // flight (one row at a time left these reductions at 0.4-0.5 of the HBM rate the elementwise passes reach)
      if (nch == 8) {
        for (; r + (OP::kBatch - 1) * stride < M; r += OP::kBatch * stride) {
          typename OP::Regs rg[OP::kBatch];
#pragma unroll
          for (int u = 0; u < OP::kBatch; ++u) op.load(r + u * stride, c8 * 8, rg[u]);
#pragma unroll
          for (int u = 0; u < OP::kBatch; ++u) op.acc(rg[u], a0, a1);
        }
      }
    }
    for (; r < M; r += stride) op.row(r, c8 * 8, nch, a0, a1);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_acc[threadIdx.x * 16 + j] = a0[j];
    s_acc[threadIdx.x * 16 + 8 + j] = a1[j];
  }
  __syncthreads();
  const int cpad = l.c8n * 8;
  for (int i = threadIdx.x; i < 2 * cpad; i += BN_THREADS) {
    const int which = i / cpad, ch = i - which * cpad;
    float s = 0.f;
    for (int q = 0; q < l.rpi; ++q) s += s_acc[(q * l.c8n + ch / 8) * 16 + which * 8 + (ch & 7)];
    part[((size_t)blockIdx.x * 2 + which) * cpad + ch] = s;
  }
}

template <typename T>
__device__ inline void load_n(const T* p, int nch, float (&v)[8]) {
  if (nch == 8) {
    load8(p, v);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < nch ? Elem<T>::ld(p + j) : 0.f;
  }
}


// out[i] = sum_p part[p][i], summed in double in a fixed order.  CTA = 32 outputs x 8 slices of the partials (lane = output, so
// every load is a coalesced 128 B row segment); the slices meet in shared memory and are added in slice order.  One thread per
// output walking all partials serially left ~10 CTAs on the machine for the gate-shift / conv weight gradients.
constexpr int PS_SLICES = 8;
static __global__ void __launch_bounds__(32 * PS_SLICES)
partial_sum_kernel(const float* __restrict__ part, int nparts, long long count, float* __restrict__ out) {
  __shared__ double s_p[PS_SLICES][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + lane;
  double s = 0.0;
  if (i < count)
    for (int p = slice; p < nparts; p += PS_SLICES) s += (double)part[(size_t)p * count + i];
  s_p[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && i < count) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < PS_SLICES; ++q) t += s_p[q][lane];
    out[i] = (float)t;
  }
}
static inline void launch_partial_sum(const float* part, int nparts, long long count, float* out, cudaStream_t st) {
  partial_sum_kernel<<<(unsigned)ceil_div_ll(count, 32), 32 * PS_SLICES, 0, st>>>(part, nparts, count, out);
}

inline int bn_grid(long long M, int C) {
  const BnLayout l = bn_layout(C);
  const long long need = ceil_div_ll(M, l.rpi);
  return (int)(need < BN_MAX_GRID ? need : BN_MAX_GRID);
}

}  // namespace tdeed
