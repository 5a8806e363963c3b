// vad.cu -- frame-wise probabilities -> sound-event (onset, offset) frame indices on the device.
//
// Replaces the per-(clip, class) Python loops of
//   frame_prediction_to_event_prediction     /root/reference/utils/utilities.py:70-123
//   activity_detection and helpers           /root/reference/utils/vad.py:11-134
// (the inner loop of optimize_sed_thresholds, /root/reference/utils/optimize_thresholds.py:66-85, which
// re-runs it for every candidate threshold set).  Integer work: results must be bit-exact.
//
// One thread per (clip n, class k) series.  The series is walked once; the reference's four list passes
// (runs -> double threshold + smooth(1) -> smooth(n_smooth) -> salt removal) are chained as streaming
// stages with O(1) state each, so nothing is materialised.  Threads of a warp own consecutive classes of the
// (N, T, K) tensor: the frame reads are coalesced.  Two launches: count (pairs per series) and, after an
// exclusive scan of the counts by the caller, fill.
//
// The reference's asymmetric index arithmetic is part of its behaviour and is kept (see oracle/vad.py).
// A non-first run that begins on the last frame makes the reference read x[T] (IndexError, vad.py:78);
// such series are reported through `flags` (bit 0) and produce no pairs.
#include "common.cuh"

namespace sed {
namespace {

struct Series {
  const float* x;   // element t at x[t * stride]
  int stride, T;
  __device__ __forceinline__ float at(int t) const { return x[(long long)t * stride]; }
};

template <bool kFill>
struct Sink {
  int count;
  int* out;        // pairs (bgn, fin)
  int n_salt;
  __device__ __forceinline__ void emit(int bgn, int fin) {     // remove_salt_noise, vad.py:122-134
    if (fin - bgn > n_salt) {
      if (kFill) { out[2 * count] = bgn; out[2 * count + 1] = fin; }
      ++count;
    }
  }
};

template <bool kFill>
__global__ void vad_kernel(const float* __restrict__ frame, const float* __restrict__ clip, int N, int T, int K,
                           const float* __restrict__ at_thres, const float* __restrict__ hi_thres,
                           const float* __restrict__ lo_thres, const int* __restrict__ n_smooth,
                           const int* __restrict__ n_salt, const long long* __restrict__ offsets,
                           int* __restrict__ counts, int* __restrict__ pairs, int* __restrict__ flags) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)N * K) return;
  const int n = (int)(idx / K), k = (int)(idx % K);
  Sink<kFill> sink;
  sink.count = 0;
  sink.n_salt = n_salt[k];
  sink.out = kFill ? pairs + 2 * offsets[idx] : nullptr;
  int flag = 0;
  if (clip == nullptr || clip[idx] > at_thres[k]) {             // utilities.py:103-104
    Series s{frame + (long long)n * T * K + k, K, T};
    const float hi = hi_thres[k];
    const bool two = lo_thres != nullptr;
    const float lo = two ? lo_thres[k] : 0.f;
    // stage state: smooth(1) after the double threshold, then smooth(n_smooth)
    int a_mem = 0, a_pre = 0; bool a_has = false;                // inner smooth, n = 1
    int b_mem = 0, b_pre = 0; bool b_has = false;                // outer smooth, n = n_smooth[k]
    const int nsm = n_smooth[k];
    auto outer_push = [&](int bgn, int fin) {
      if (!b_has) { b_mem = bgn; b_has = true; }
      else if (bgn - b_pre > nsm) { sink.emit(b_mem, b_pre); b_mem = bgn; }
      b_pre = fin;
    };
    auto inner_push = [&](int bgn, int fin) {
      if (!a_has) { a_mem = bgn; a_has = true; }
      else if (bgn - a_pre > 1) { outer_push(a_mem, a_pre); a_mem = bgn; }
      a_pre = fin;
    };
    auto stage_push = [&](int bgn, int fin) {
      if (!two) { outer_push(bgn, fin); return; }
      if (bgn >= T) { flag = 1; return; }                        // the reference raises IndexError here
      while (bgn != -1) {                                        // vad.py:77-80
        if (s.at(bgn) < lo) break;
        --bgn;
      }
      while (fin != T) {                                         // vad.py:82-85
        if (s.at(fin) < lo) break;
        ++fin;
      }
      inner_push(bgn + 1, fin);
    };
    // runs above the high threshold -> the reference's [bgn, fin] pairs (vad.py:44-66)
    int run = 0, rs = -1, re = -1;
    for (int t = 0; t < T; ++t) {
      if (s.at(t) > hi) {
        if (rs >= 0 && re == t - 1) {
          re = t;
        } else {
          if (rs >= 0) { stage_push(run == 0 ? rs : rs + 1, re + 1); ++run; }
          rs = re = t;
        }
      }
    }
    if (rs >= 0) stage_push(run == 0 ? rs : rs + 1, re);
    if (two && a_has) outer_push(a_mem, a_pre);
    if (b_has) sink.emit(b_mem, b_pre);
    if (flag) sink.count = 0;
  }
  if (!kFill) {
    counts[idx] = sink.count;
    if (flags) flags[idx] = flag;
  }
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_vad_count(const float* frame, const float* clip, int N, int T, int K, const float* at_thres,
                  const float* hi_thres, const float* lo_thres, const int* n_smooth, const int* n_salt,
                  int* counts, int* flags, sed_stream_t stream) {
  SED_REQUIRE(frame && hi_thres && n_smooth && n_salt && counts, "sed_vad_count: null pointer");
  SED_REQUIRE(clip == nullptr || at_thres != nullptr, "sed_vad_count: clip-wise gating needs at_thres");
  SED_REQUIRE(N >= 0 && T >= 0 && K >= 1, "sed_vad_count: bad shape");
  if (N == 0) return 0;
  const long long total = (long long)N * K;
  vad_kernel<false><<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      frame, clip, N, T, K, at_thres, hi_thres, lo_thres, n_smooth, n_salt, nullptr, counts, nullptr, flags);
  SED_LAUNCH_CHECK("vad_kernel<count>");
  return 0;
}

int sed_vad_fill(const float* frame, const float* clip, int N, int T, int K, const float* at_thres,
                 const float* hi_thres, const float* lo_thres, const int* n_smooth, const int* n_salt,
                 const long long* offsets, int* pairs, sed_stream_t stream) {
  SED_REQUIRE(frame && hi_thres && n_smooth && n_salt && offsets && pairs, "sed_vad_fill: null pointer");
  SED_REQUIRE(clip == nullptr || at_thres != nullptr, "sed_vad_fill: clip-wise gating needs at_thres");
  SED_REQUIRE(N >= 0 && T >= 0 && K >= 1, "sed_vad_fill: bad shape");
  if (N == 0) return 0;
  const long long total = (long long)N * K;
  vad_kernel<true><<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      frame, clip, N, T, K, at_thres, hi_thres, lo_thres, n_smooth, n_salt, offsets, nullptr, pairs, nullptr);
  SED_LAUNCH_CHECK("vad_kernel<fill>");
  return 0;
}

}  // extern "C"
