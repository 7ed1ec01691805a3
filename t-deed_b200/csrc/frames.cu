// Frame-feature cache plumbing of the video-level engine (tdeed_gather_rows, include/tdeed_b200.h (12)).
//
// The reference runs every overlapping clip of a video through the whole network (dataset/frame.py:409-423 builds
// clips with 75 % overlap, util/eval.py:289-349 feeds them one by one), so each frame passes the clip-independent
// part of the backbone (stem, s1, s2 — everything before the first GatedShift, model/shift.py:47-59) four times.
// Here those features are computed once per unique frame into a ring in HBM; this kernel assembles the clip batches
// (and files new frames into the ring): a pure row copy with index lists, 16 B per thread per trip, HBM bound.
#include "common.cuh"

namespace tdeed {

constexpr int GR_THREADS = 256;

// one CTA row-slice: blockIdx.x = destination row entry, blockIdx.y = slice of the row
__global__ void __launch_bounds__(GR_THREADS)
gather_rows_kernel(const uint4* __restrict__ src, const uint4* __restrict__ pad_row, uint4* __restrict__ dst,
                   const int* __restrict__ src_idx, const int* __restrict__ dst_idx, long long row_vec, int slices) {
  const long long i = blockIdx.x;
  const int s = src_idx[i];
  const long long d = dst_idx ? (long long)dst_idx[i] : i;
  const uint4* from = s < 0 ? pad_row : src + (long long)s * row_vec;
  uint4* to = dst + d * row_vec;
  const long long per = ceil_div_ll(row_vec, slices);
  const long long lo = per * blockIdx.y, hi = lo + per < row_vec ? lo + per : row_vec;
  for (long long v = lo + threadIdx.x; v < hi; v += GR_THREADS) to[v] = __ldg(from + v);
}

}  // namespace tdeed

extern "C" int tdeed_gather_rows(const void* src, const void* pad_row, void* dst, const int* src_idx, const int* dst_idx,
                                 int n_rows, long long row_bytes, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(src && dst && src_idx, TDEED_ERR_SHAPE, "tdeed_gather_rows: null pointer");
  TDEED_REQUIRE(n_rows > 0 && row_bytes > 0 && row_bytes % 16 == 0, TDEED_ERR_SHAPE,
                "tdeed_gather_rows: n_rows=%d row_bytes=%lld (must be a positive multiple of 16)", n_rows, row_bytes);
  TDEED_REQUIRE(((uintptr_t)src | (uintptr_t)dst | (uintptr_t)pad_row) % 16 == 0, TDEED_ERR_SHAPE,
                "tdeed_gather_rows: pointers must be 16-byte aligned");
  const long long row_vec = row_bytes / 16;
  // enough CTAs to fill the machine even for a handful of rows: slices of >= 4 KB
  long long slices = ceil_div_ll(row_vec, 256);
  const long long want = ceil_div_ll(4LL * kNumSMs, n_rows);
  if (slices > want) slices = want < 1 ? 1 : want;
  if (slices > 65535) slices = 65535;
  dim3 grid((unsigned)n_rows, (unsigned)slices);
  gather_rows_kernel<<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>((const uint4*)src, (const uint4*)pad_row, (uint4*)dst, src_idx,
                                                                     dst_idx, row_vec, (int)slices);
  return check_launch("tdeed_gather_rows");
}
