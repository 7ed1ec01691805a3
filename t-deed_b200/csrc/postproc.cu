// (11) per-video post-processing: clip -> video accumulation, event extraction, NMS / soft-NMS.
// Reference: util/eval.py:284-349 (accumulate), :87-193 (process_frame_predictions[_challenge]),
// :195-227 (non_maximum_supression), :229-261 (soft_non_maximum_supression).
// Everything here is integer / IEEE-exact fp32 / fp64 arithmetic in the reference's operation order,
// so results are bit-identical to the Python code (asserted in tests/test_postproc_gpu.py).
#include "common.cuh"

namespace tdeed {

// numpy's pairwise summation of a contiguous float32 vector with n < 128 elements (the row sums of
// `pred_scores.sum(axis=1)`, util/eval.py:317): 8 running accumulators, fixed combine tree, scalar tail.
__device__ inline float numpy_rowsum(const float* __restrict__ p, int n) {
  if (n < 8) {
    float s = 0.f;      // numpy starts from the first element; 0 + p0 is exact
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, p[i]);
    return s;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = p[j];
  int i = 8;
  for (; i + 8 <= n; i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], p[i + j]);
  }
  float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                      __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) s = __fadd_rn(s, p[i]);
  return s;
}

__global__ void __launch_bounds__(256)
clip_accumulate_kernel(float* __restrict__ scores, int* __restrict__ support, int video_len, int K,
                       const float* __restrict__ pred, const int* __restrict__ starts, int n_clips, int T, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)video_len * K) return;
  const int l = (int)(idx / K), k = (int)(idx - (long long)l * K);
  float s = scores[idx];
  int sup = 0;
  bool touched = false;
  for (int i = 0; i < n_clips; ++i) {          // clip order == the reference's `+=` order
    const int t = l - starts[i];
    if (t < 0 || t >= T) continue;
    const float* row = pred + ((size_t)i * T + t) * K;
    s = __fadd_rn(s, row[k]);
    touched = true;
    if (k == 0) sup += (mode == 0) ? (numpy_rowsum(row, K) != 0.f ? 1 : 0) : 1;
  }
  if (touched) {
    scores[idx] = s;
    if (k == 0 && sup) support[l] += sup;
  }
}

// same, with the clip starts passed by value (no index upload: see csrc/frames.cu)
struct ClipStarts { int v[TDEED_MAX_STARTS_PER_CALL]; };

__global__ void __launch_bounds__(256)
clip_accumulate_hs_kernel(float* __restrict__ scores, int* __restrict__ support, int video_len, int K,
                          const float* __restrict__ pred, const ClipStarts starts, int n_clips, int T, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)video_len * K) return;
  const int l = (int)(idx / K), k = (int)(idx - (long long)l * K);
  float s = scores[idx];
  int sup = 0;
  bool touched = false;
  for (int i = 0; i < n_clips; ++i) {          // clip order == the reference's `+=` order
    const int t = l - starts.v[i];
    if (t < 0 || t >= T) continue;
    const float* row = pred + ((size_t)i * T + t) * K;
    s = __fadd_rn(s, row[k]);
    touched = true;
    if (k == 0) sup += (mode == 0) ? (numpy_rowsum(row, K) != 0.f ? 1 : 0) : 1;
  }
  if (touched) {
    scores[idx] = s;
    if (k == 0 && sup) support[l] += sup;
  }
}

// ---- extraction: one CTA per video, frames processed in order in chunks of blockDim ----
constexpr int EX_THREADS = 1024;

__device__ inline int block_excl_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    s_warp[lane] = winc - w;           // exclusive warp offsets
    if (lane == 31) s_warp[32] = winc; // block total
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[warp] + inc - v;
}

__global__ void __launch_bounds__(EX_THREADS)
extract_events_kernel(float* __restrict__ scores, int* __restrict__ support, int video_len, int K, float thr,
                      int* __restrict__ pred, int* __restrict__ ev_frame, int* __restrict__ ev_label,
                      float* __restrict__ ev_score, int* __restrict__ hr_frame, int* __restrict__ hr_label,
                      float* __restrict__ hr_score, int* __restrict__ counts) {
  __shared__ int s_warp[33];
  int ev_base = 0, hr_base = 0;
  for (int base = 0; base < video_len; base += EX_THREADS) {
    const int l = base + threadIdx.x;
    int is_ev = 0, n_hr = 0, arg = 0;
    float best = 0.f;
    if (l < video_len) {
      int sup = support[l];
      if (sup == 0) { sup = 1; support[l] = 1; }
      const float d = (float)sup;
      float* row = scores + (size_t)l * K;
      for (int k = 0; k < K; ++k) {
        const float v = __fdiv_rn(row[k], d);
        row[k] = v;
        if (k == 0 || v > best) { best = v; arg = k; }   // first maximum, like np.argmax
        if (k >= 1 && v >= thr) ++n_hr;
      }
      pred[l] = arg;
      is_ev = arg != 0;
    }
    int ev_tot, hr_tot;
    const int ev_off = block_excl_scan(is_ev, s_warp, &ev_tot);
    const int hr_off = block_excl_scan(n_hr, s_warp, &hr_tot);
    if (l < video_len) {
      if (is_ev) {
        ev_frame[ev_base + ev_off] = l;
        ev_label[ev_base + ev_off] = arg;
        ev_score[ev_base + ev_off] = best;
      }
      int o = hr_base + hr_off;
      const float* row = scores + (size_t)l * K;
      for (int k = 1; k < K; ++k) {
        const float v = row[k];
        if (v >= thr) { hr_frame[o] = l; hr_label[o] = k; hr_score[o] = v; ++o; }
      }
    }
    ev_base += ev_tot;
    hr_base += hr_tot;
  }
  if (threadIdx.x == 0) { counts[0] = ev_base; counts[1] = hr_base; }
}

// ---- NMS ----
constexpr int NMS_THREADS = 512;
constexpr int NMS_MAXK = 64;

struct NmsHeader {            // lives at the start of the workspace
  int hist[NMS_MAXK];
  int first[NMS_MAXK];        // index of first appearance (label-bucket order of the reference)
  int offs[NMS_MAXK + 1];
  int sel_cnt[NMS_MAXK];
  int sel_offs[NMS_MAXK + 1];
};

struct NmsWs {
  NmsHeader* hdr;
  int* seg_frame;
  int* seg_flag;
  double* seg_score;
  int* sel_frame;
  double* sel_score;
};

__host__ __device__ inline NmsWs nms_ws(void* base, int capacity) {
  NmsWs w;
  char* p = reinterpret_cast<char*>(base);
  w.hdr = reinterpret_cast<NmsHeader*>(p);
  p += 4096;
  w.seg_score = reinterpret_cast<double*>(p); p += (size_t)capacity * 8;
  w.sel_score = reinterpret_cast<double*>(p); p += (size_t)capacity * 8;
  w.seg_frame = reinterpret_cast<int*>(p); p += (size_t)capacity * 4;
  w.seg_flag = reinterpret_cast<int*>(p); p += (size_t)capacity * 4;
  w.sel_frame = reinterpret_cast<int*>(p);
  return w;
}

__global__ void __launch_bounds__(NMS_THREADS)
nms_hist_kernel(const int* __restrict__ label, const int* __restrict__ n_dev, int capacity, int K, void* wsbase) {
  __shared__ int s_hist[NMS_MAXK], s_first[NMS_MAXK];
  NmsWs ws = nms_ws(wsbase, capacity);
  const int n = min(*n_dev, capacity);
  for (int i = threadIdx.x; i < NMS_MAXK; i += NMS_THREADS) { s_hist[i] = 0; s_first[i] = 0x7fffffff; }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
    const int l = label[i];
    if (l >= 0 && l < K) { atomicAdd(&s_hist[l], 1); atomicMin(&s_first[l], i); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = 0;
    for (int l = 0; l < NMS_MAXK; ++l) {
      ws.hdr->hist[l] = s_hist[l];
      ws.hdr->first[l] = s_first[l];
      ws.hdr->offs[l] = o;
      o += s_hist[l];
    }
    ws.hdr->offs[NMS_MAXK] = o;
  }
}

// priority: higher score first, ties -> lower list index first (max() returns the first maximum)
__device__ inline bool outranks(double sj, int j, double si, int i) { return sj > si || (sj == si && j < i); }

// flag bits
constexpr int F_ALIVE = 1, F_SEL = 2, F_DONE = 4;

__global__ void __launch_bounds__(NMS_THREADS)
nms_segment_kernel(const int* __restrict__ frame, const int* __restrict__ label, const float* __restrict__ score,
                   const int* __restrict__ n_dev, int capacity, int window, double thr, int soft, void* wsbase) {
  __shared__ int s_warp[33];
  NmsWs ws = nms_ws(wsbase, capacity);
  const int lab = blockIdx.x + 1;
  const int n = min(*n_dev, capacity);
  const int m = ws.hdr->hist[lab];
  const int off = ws.hdr->offs[lab];
  int* sf = ws.seg_frame + off;
  int* fl = ws.seg_flag + off;
  double* ss = ws.seg_score + off;

  // stable compaction of this label's events (keeps frame order)
  int base = 0;
  for (int b0 = 0; b0 < n; b0 += NMS_THREADS) {
    const int i = b0 + threadIdx.x;
    const int mine = (i < n && label[i] == lab) ? 1 : 0;
    int tot;
    const int o = block_excl_scan(mine, s_warp, &tot);
    if (mine) {
      sf[base + o] = frame[i];
      ss[base + o] = (double)score[i];
      fl[base + o] = F_ALIVE;
    }
    base += tot;
  }
  __syncthreads();

  // rounds: every alive candidate >= thr that outranks all alive neighbours within +-sel_window is selected
  // simultaneously, then its +-window neighbourhood is suppressed (hard) or decayed (soft).
  // hard: sel_window = window (selected events are > window apart; suppression is order independent).
  // soft: sel_window = 2*window, so two events selected in one round never share a decayed neighbour and
  //       an event selected in a later round was outranked by every earlier one within 2*window -> the
  //       fp64 decays hit each score in exactly the reference's sequential (descending-score) order.
  const int sel_window = soft ? 2 * window : window;
  while (true) {
    int any = 0;
    for (int i = threadIdx.x; i < m; i += NMS_THREADS) {
      if (!(fl[i] & F_ALIVE)) continue;
      const double si = ss[i];
      if (si < thr) continue;
      any = 1;
      const int fi = sf[i];
      bool top = true;
      for (int j = i - 1; j >= 0 && fi - sf[j] <= sel_window && top; --j)
        if ((fl[j] & F_ALIVE) && outranks(ss[j], j, si, i)) top = false;
      for (int j = i + 1; j < m && sf[j] - fi <= sel_window && top; ++j)
        if ((fl[j] & F_ALIVE) && outranks(ss[j], j, si, i)) top = false;
      if (top) fl[i] |= F_SEL;
    }
    if (!__syncthreads_or(any)) break;
    // apply: read-only on F_SEL of neighbours, writes only own entry
    for (int i = threadIdx.x; i < m; i += NMS_THREADS) {
      const int f = fl[i];
      if (!(f & F_ALIVE) || (f & F_SEL)) continue;
      const int fi = sf[i];
      int jl = -1, jr = -1;
      for (int j = i - 1; j >= 0 && fi - sf[j] <= window; --j)
        if (fl[j] & F_SEL) { jl = j; break; }
      for (int j = i + 1; j < m && sf[j] - fi <= window; ++j)
        if (fl[j] & F_SEL) { jr = j; break; }
      if (jl < 0 && jr < 0) continue;
      if (!soft) {
        fl[i] = 0;                                  // suppressed
      } else {
        int j1 = jl, j2 = jr;                       // apply decays in emission order (higher priority first)
        if (jl >= 0 && jr >= 0 && outranks(ss[jr], jr, ss[jl], jl)) { j1 = jr; j2 = jl; }
        if (j1 < 0) { j1 = j2; j2 = -1; }
        double s = ss[i];
        const double w2 = (double)((long long)window * window);
        {
          const long long d = llabs((long long)sf[j1] - fi);
          s = __ddiv_rn(__dmul_rn(s, (double)(d * d)), w2);
        }
        if (j2 >= 0) {
          const long long d = llabs((long long)sf[j2] - fi);
          s = __ddiv_rn(__dmul_rn(s, (double)(d * d)), w2);
        }
        ss[i] = s;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += NMS_THREADS)
      if (fl[i] & F_SEL) fl[i] = F_DONE;            // emitted with its current score, removed from the list
    __syncthreads();
  }

  // compact the selected events of this label (frame order)
  int* of = ws.sel_frame + off;
  double* os = ws.sel_score + off;
  base = 0;
  for (int b0 = 0; b0 < m; b0 += NMS_THREADS) {
    const int i = b0 + threadIdx.x;
    const int mine = (i < m && (fl[i] & F_DONE)) ? 1 : 0;
    int tot;
    const int o = block_excl_scan(mine, s_warp, &tot);
    if (mine) { of[base + o] = sf[i]; os[base + o] = ss[i]; }
    base += tot;
  }
  if (threadIdx.x == 0) ws.hdr->sel_cnt[lab] = base;
}

__device__ inline int lower_bound_i(const int* a, int n, int v) {   // first index with a[i] >= v
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

// merge the per-label results into one list sorted by (frame, label-bucket order)
__global__ void __launch_bounds__(256)
nms_merge_kernel(int capacity, int K, void* wsbase, int* __restrict__ out_frame, int* __restrict__ out_label,
                 double* __restrict__ out_score, int* __restrict__ out_count) {
  NmsWs ws = nms_ws(wsbase, capacity);
  const NmsHeader* h = ws.hdr;
  int total = 0;
  for (int l = 1; l < K; ++l) total += h->sel_cnt[l];
  if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = total;
  for (int lab = 1; lab < K; ++lab) {
    const int cnt = h->sel_cnt[lab], off = h->offs[lab];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
      const int f = ws.sel_frame[off + i];
      int pos = i;
      for (int l2 = 1; l2 < K; ++l2) {
        if (l2 == lab || h->sel_cnt[l2] == 0) continue;
        const bool before = h->first[l2] < h->first[lab];     // bucket l2 precedes bucket lab on equal frames
        pos += lower_bound_i(ws.sel_frame + h->offs[l2], h->sel_cnt[l2], before ? f + 1 : f);
      }
      out_frame[pos] = f;
      out_label[pos] = lab;
      out_score[pos] = ws.sel_score[off + i];
    }
  }
}

}  // namespace tdeed

extern "C" int tdeed_clip_accumulate(float* scores, int* support, int video_len, int K, const float* pred,
                                     const int* starts, int n_clips, int T, int mode, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(scores && support && pred && starts, TDEED_ERR_SHAPE, "tdeed_clip_accumulate: null pointer");
  TDEED_REQUIRE(video_len > 0 && K > 0 && K < 128 && n_clips > 0 && T > 0 && (mode == 0 || mode == 1), TDEED_ERR_SHAPE,
                "tdeed_clip_accumulate: bad shape L=%d K=%d clips=%d T=%d mode=%d", video_len, K, n_clips, T, mode);
  const long long total = (long long)video_len * K;
  clip_accumulate_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(scores, support, video_len, K, pred,
                                                                                             starts, n_clips, T, mode);
  return check_launch("tdeed_clip_accumulate");
}

extern "C" int tdeed_clip_accumulate_host(float* scores, int* support, int video_len, int K, const float* pred,
                                          const int* starts_host, int n_clips, int T, int mode, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(scores && support && pred && starts_host, TDEED_ERR_SHAPE, "tdeed_clip_accumulate_host: null pointer");
  TDEED_REQUIRE(video_len > 0 && K > 0 && K < 128 && n_clips > 0 && n_clips <= TDEED_MAX_STARTS_PER_CALL && T > 0 &&
                (mode == 0 || mode == 1), TDEED_ERR_SHAPE, "tdeed_clip_accumulate_host: bad shape L=%d K=%d clips=%d (max %d) T=%d mode=%d",
                video_len, K, n_clips, TDEED_MAX_STARTS_PER_CALL, T, mode);
  ClipStarts st;
  for (int i = 0; i < n_clips; ++i) st.v[i] = starts_host[i];
  const long long total = (long long)video_len * K;
  clip_accumulate_hs_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(scores, support, video_len, K, pred,
                                                                                                st, n_clips, T, mode);
  return check_launch("tdeed_clip_accumulate_host");
}

extern "C" int tdeed_extract_events(float* scores, int* support, int video_len, int K, float threshold, int* pred,
                                    int* ev_frame, int* ev_label, float* ev_score, int* hr_frame, int* hr_label,
                                    float* hr_score, int* counts_out, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(scores && support && pred && ev_frame && ev_label && ev_score && hr_frame && hr_label && hr_score && counts_out,
                TDEED_ERR_SHAPE, "tdeed_extract_events: null pointer");
  TDEED_REQUIRE(video_len > 0 && K > 1, TDEED_ERR_SHAPE, "tdeed_extract_events: bad shape L=%d K=%d", video_len, K);
  extract_events_kernel<<<1, EX_THREADS, 0, (cudaStream_t)stream>>>(scores, support, video_len, K, threshold, pred, ev_frame,
                                                                   ev_label, ev_score, hr_frame, hr_label, hr_score, counts_out);
  return check_launch("tdeed_extract_events");
}

extern "C" long long tdeed_nms_workspace_bytes(int capacity, int K) {
  (void)K;
  return 4096 + (long long)capacity * 28 + 64;
}

extern "C" int tdeed_nms(const int* frame, const int* label, const float* score, const int* n_events_dev, int capacity,
                         int K, int window, double threshold, int soft, void* workspace, int* out_frame, int* out_label,
                         double* out_score, int* out_count, void* stream) {
  using namespace tdeed;
  TDEED_REQUIRE(frame && label && score && n_events_dev && workspace && out_frame && out_label && out_score && out_count,
                TDEED_ERR_SHAPE, "tdeed_nms: null pointer");
  TDEED_REQUIRE(capacity > 0 && K > 1 && K <= NMS_MAXK && window >= 0, TDEED_ERR_SHAPE, "tdeed_nms: bad shape capacity=%d K=%d window=%d",
                capacity, K, window);
  TDEED_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, TDEED_ERR_SHAPE, "tdeed_nms: workspace must be 8-byte aligned");
  static_assert(sizeof(NmsHeader) <= 4096, "header too large");
  cudaStream_t st = (cudaStream_t)stream;
  nms_hist_kernel<<<1, NMS_THREADS, 0, st>>>(label, n_events_dev, capacity, K, workspace);
  int rc = check_launch("tdeed_nms(hist)");
  if (rc) return rc;
  nms_segment_kernel<<<K - 1, NMS_THREADS, 0, st>>>(frame, label, score, n_events_dev, capacity, window, threshold, soft, workspace);
  rc = check_launch("tdeed_nms(segments)");
  if (rc) return rc;
  nms_merge_kernel<<<64, 256, 0, st>>>(capacity, K, workspace, out_frame, out_label, out_score, out_count);
  return check_launch("tdeed_nms(merge)");
}
