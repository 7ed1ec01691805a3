// Greedy prediction <-> ground-truth matching of the mAP scorer (tdeed_match_events, include/tdeed_b200.h (13)).
//
// util/score.py:45-89 of the reference walks the predictions of one class in descending-score order and, for each, scans
// the ground-truth frames of the prediction's video for the closest one not recalled yet; a hit within the tolerance
// marks it recalled.  The recalled set is keyed by (video, frame), so matching never crosses videos: every
// (class, video) segment is an independent sequential problem, and every tolerance is another independent copy.
// One warp per (segment, tolerance): lanes scan the segment's ground truth, a (distance, index) min-reduction picks the
// reference's choice (strictly-closer replacement == lowest list index among equal distances).  Integer work, bit-exact.
#include "common.cuh"

namespace tdeed {

constexpr int ME_WARPS = 4;

__global__ void __launch_bounds__(ME_WARPS * 32)
match_events_kernel(const int* __restrict__ pred_frame, const int* __restrict__ pred_off, const int* __restrict__ gt_frame,
                    const int* __restrict__ gt_off, int n_units, int total_pred, int total_gt,
                    const int* __restrict__ tolerances, unsigned char* __restrict__ recalled, unsigned char* __restrict__ tp) {
  const int unit = blockIdx.x * ME_WARPS + (threadIdx.x >> 5);
  if (unit >= n_units) return;
  const int lane = threadIdx.x & 31;
  const int tol = tolerances[blockIdx.y];
  const int p0 = pred_off[unit], p1 = pred_off[unit + 1], g0 = gt_off[unit], g1 = gt_off[unit + 1];
  unsigned char* rec = recalled + (size_t)blockIdx.y * total_gt;
  unsigned char* out = tp + (size_t)blockIdx.y * total_pred;
  for (int g = g0 + lane; g < g1; g += 32) rec[g] = 0;
  __syncwarp();
  for (int p = p0; p < p1; ++p) {
    const int f = pred_frame[p];
    unsigned long long best = ~0ull;                      // (distance << 32) | list index
    for (int g = g0 + lane; g < g1; g += 32) {
      if (rec[g]) continue;
      const int d = abs(f - gt_frame[g]);
      const unsigned long long key = ((unsigned long long)(unsigned)d << 32) | (unsigned)(g - g0);
      best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    const bool hit = best != ~0ull && (int)(best >> 32) <= tol;
    if (hit) {
      // the reference's `recalled` is a set of (video, frame) VALUES: duplicate ground-truth frames fall together
      const int gf = gt_frame[g0 + (int)(best & 0xffffffffu)];
      for (int g = g0 + lane; g < g1; g += 32)
        if (gt_frame[g] == gf) rec[g] = 1;
    }
    if (lane == 0) out[p] = hit ? 1 : 0;
    __syncwarp();
  }
}

}  // namespace tdeed

extern "C" int tdeed_match_events(const int* pred_frame, const int* pred_off, const int* gt_frame, const int* gt_off, int n_units,
                                  int total_pred, int total_gt, const int* tolerances, int n_tol, unsigned char* recalled_ws,
                                  unsigned char* tp, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(pred_off && gt_off && tolerances && tp, TDEED_ERR_SHAPE, "tdeed_match_events: null pointer");
  TDEED_REQUIRE(n_units > 0 && n_tol > 0 && n_tol <= 65535 && total_pred >= 0 && total_gt >= 0, TDEED_ERR_SHAPE,
                "tdeed_match_events: n_units=%d n_tol=%d total_pred=%d total_gt=%d", n_units, n_tol, total_pred, total_gt);
  TDEED_REQUIRE((total_pred == 0 || pred_frame) && (total_gt == 0 || (gt_frame && recalled_ws)), TDEED_ERR_SHAPE,
                "tdeed_match_events: null event arrays");
  dim3 grid((unsigned)ceil_div(n_units, ME_WARPS), (unsigned)n_tol);
  match_events_kernel<<<grid, ME_WARPS * 32, 0, (cudaStream_t)stream>>>(pred_frame, pred_off, gt_frame, gt_off, n_units, total_pred,
                                                                         total_gt, tolerances, recalled_ws, tp);
  return check_launch("tdeed_match_events");
}
