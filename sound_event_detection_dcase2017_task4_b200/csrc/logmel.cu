// logmel.cu -- fused STFT(1024, Hann, reflect) -> |.|^2 -> sparse mel -> 10*log10 front-end.
//
// Replaces torchlibrosa.stft.Spectrogram + LogmelFilterBank as called at
// /root/reference/pytorch/models.py:199-200 (ctor contracts :166-173).  The reference evaluates
// the windowed DFT as two dense fp32 conv1d (2.1 GFLOP per 10 s clip) and materialises the
// (T,513) power spectrogram; here every frame is a real 1024-point FFT done by ONE WARP in registers
// (512-point complex FFT of z[n] = x[2n] + i x[2n+1] + the even/odd split, see logmel_warp_kernel) and only the
// (T,64) log-mel leaves the SM.
#include "common.cuh"

namespace sed {
namespace {

constexpr int kNfft = 1024;
constexpr int kBins = kNfft / 2 + 1;   // 513
constexpr int kHalf = kNfft / 2;       // 512 complex points
constexpr int kFramesPerChunk = 16;    // consecutive frames of a clip per work item (upper bound)
// Mel projection schedule (built once per CTA in shared memory): every mel filter is cut into pieces of at most
// kPieceTaps consecutive taps; piece p is owned by lane p % 32 of the frame's warp.  The reference bank has 866 taps in
// filters of 3..47 taps: one thread per filter made the 47-tap thread the critical path and every weight a scattered
// 4-byte global load (ncu r01: 78 % excessive sectors, L1/TEX 80 % busy).  Pieces balance the lanes (<= 34 taps per
// lane, 27 on average) and the weights are staged tap-major / piece-minor so a warp reads them without conflicts.
constexpr int kPieceTaps = 16;
constexpr int kMaxPieces = 128;
constexpr int kMaxMels = 128;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// complex add / sub on Blackwell's packed fp32 pipe (FADD2: two IEEE adds per issue slot; the butterflies are
// add-dominated and the kernel is issue-limited between its shared-memory exchanges)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// exp(-2*pi*i * num / den)
__device__ __forceinline__ float2 twiddle(int num, int den) {
  float s, c;
  sincospif(2.0f * (float)num / (float)den, &s, &c);
  return make_float2(c, -s);
}

// 4-point forward DFT (W4 = -i), natural order out
__device__ __forceinline__ void dft4(float2 c0, float2 c1, float2 c2, float2 c3, float2& o0,
                                     float2& o1, float2& o2, float2& o3) {
  float2 s0 = cadd(c0, c2), s1 = csub(c0, c2), s2 = cadd(c1, c3), s3 = mul_neg_i(csub(c1, c3));
  o0 = cadd(s0, s2);
  o2 = csub(s0, s2);
  o1 = cadd(s1, s3);
  o3 = csub(s1, s3);
}

// in-place 8-point forward DFT, natural order:  v[q] <- sum_r v[r] * W8^(r q)
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  // b_r *= W8^r :  W8 = (1 - i)/sqrt2,  W8^2 = -i,  W8^3 = (-1 - i)/sqrt2
  b1 = make_float2(h * (b1.x + b1.y), h * (b1.y - b1.x));
  b2 = mul_neg_i(b2);
  b3 = make_float2(h * (b3.y - b3.x), -h * (b3.x + b3.y));
  dft4(a0, a1, a2, a3, v[0], v[2], v[4], v[6]);
  dft4(b0, b1, b2, b3, v[1], v[3], v[5], v[7]);
}

template <typename T> struct Sample;
template <> struct Sample<float> {
  static __device__ __forceinline__ float cvt(float v) { return v; }
};
template <> struct Sample<int16_t> {
  // utils/utilities.py:66-67: (x / 32767.) in float64 then cast to float32, i.e. the correctly rounded quotient.
  // An IEEE fp32 division is ~15 issue slots per sample in a kernel that is issue-bound (the int16 kernel ran 2.4x
  // slower than the fp32 one).  Here: v -> float through the 1.5 * 2^23 magic constant (integer add + float subtract,
  // no I2F), then q0 = v * r, e = v - q0 * 32767 (exact in one FMA), q = q0 + e * r with r = fl(1 / 32767): one Newton
  // correction of the quotient, which is the correctly rounded result for every int16 value
  // (tests/test_host_logic.py proves all 65536 in exact rational arithmetic; tests/test_gpu_logmel.py checks the device).
  static __device__ __forceinline__ float cvt(int16_t v) {
    const float f = __int_as_float(0x4B400000 + (int)v) - 12582912.0f;
    const float r = 1.0f / 32767.0f;
    const float q0 = f * r;
    const float e = fmaf(-q0, 32767.0f, f);
    return fmaf(e, r, q0);
  }
};

struct MelBank {
  const float* w;
  const int* lo;
  const int* off;
  int n_mels;
  float amin, db_offset;
};

// 16-point forward DFT in registers, natural order in and out: two DFT8 (even / odd inputs) + one radix-2 stage
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  const float h = 0.70710678118654752440f, c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
  float2 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = v[2 * i]; b[i] = v[2 * i + 1]; }
  dft8(a);
  dft8(b);
  // b[q] *= W16^q,  W16 = exp(-i pi / 8)
  b[1] = cmul(b[1], make_float2(c1, -s1));
  b[2] = make_float2(h * (b[2].x + b[2].y), h * (b[2].y - b[2].x));
  b[3] = cmul(b[3], make_float2(s1, -c1));
  b[4] = mul_neg_i(b[4]);
  b[5] = cmul(b[5], make_float2(-s1, -c1));
  b[6] = make_float2(h * (b[6].y - b[6].x), -h * (b[6].x + b[6].y));
  b[7] = cmul(b[7], make_float2(-c1, -s1));
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    v[q] = cadd(a[q], b[q]);
    v[q + 8] = csub(a[q], b[q]);
  }
}
__device__ __forceinline__ float2 shfl2(float2 v, int src) {
  return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
__device__ __forceinline__ float2 shfl2_xor(float2 v, int m) {
  return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ float2 sel2(bool c, float2 a, float2 b) { return make_float2(c ? a.x : b.x, c ? a.y : b.y); }

// ---------------------------------------------------------------------------------------------------------------
// One frame per WARP (round 2).  The round-1 kernel gave a frame to 64 threads and exchanged data through shared
// memory five times per frame (three radix-8 passes, the real-FFT split, the power spectrum): 459 shared-memory
// wavefronts and ~1680 warp instructions per frame, 28 % of them address arithmetic, with five named barriers
// (ncu, profiles/r02_logmel_ncu.md).  Here lane L of a warp holds 16 complex values and the 512-point complex FFT is
//   z[n], n = L + 32 m            lane L, slot m        (coalesced 8-byte loads, window in registers)
//   DFT16 over m                  -> E_h[q]             in registers              (r = L & 15, h = L >> 4)
//   radix-2 across lanes L, L^16  -> Zr[kj]             one shuffle exchange: the 32-point DFT over n = r + 16 j'
//   transpose T[r][kj]            -> lane kj, slot r    the ONLY shared-memory exchange of the transform
//   twiddle W512^(r kj), DFT16 over r -> Z[kj + 32 kr'] in registers
// and the real-FFT split pairs bin k = kj + 32 kr' with 512 - k, which lives in lane 32 - kj, slot 15 - kr': a fixed
// lane permutation, i.e. one more shuffle exchange (lane 0 pairs with itself: its slots are rotated by one first).
// Nothing but __syncwarp() orders the frame; a CTA is four independent warps walking consecutive frames so that the
// 3.2x overlap of neighbouring frames is served by L1.
constexpr int kWarpsW = 4;
constexpr int kThreadsW = kWarpsW * 32;
constexpr int kTS = 33;                      // float2 stride of the transpose rows (odd: conflict-free both ways)
constexpr int kPLen = 520;                   // 513 power bins + padding

template <typename InT>
__device__ __forceinline__ float2 load_pair(const InT* __restrict__ src, long long i);
template <>
__device__ __forceinline__ float2 load_pair<float>(const float* __restrict__ src, long long i) {
  return __ldg(reinterpret_cast<const float2*>(src + i));
}
template <>
__device__ __forceinline__ float2 load_pair<int16_t>(const int16_t* __restrict__ src, long long i) {
  const int raw = __ldg(reinterpret_cast<const int*>(src + i));
  return make_float2(Sample<int16_t>::cvt((int16_t)(raw & 0xFFFF)), Sample<int16_t>::cvt((int16_t)(raw >> 16)));
}

// kPower: write the 513-bin power spectrogram instead of the log-mel (unfused seam A).
// Window in shared memory (one table per CTA, 16 conflict-free 8-byte loads per frame) instead of 32 registers per lane:
// 128 registers -> four CTAs (16 warps) per SM instead of three.  Measured 597 vs 575 Mframes/s at 512 x 10 s clips.
#ifndef SED_LOGMEL_SMEM_WIN
#define SED_LOGMEL_SMEM_WIN 1
#endif
#ifndef SED_LOGMEL_CTAS
#define SED_LOGMEL_CTAS 4
#endif
template <typename InT, bool kPower>
__global__ void __launch_bounds__(kThreadsW, SED_LOGMEL_CTAS)
logmel_warp_kernel(const InT* __restrict__ wave, int n_clips, int n_samples, int hop, int n_frames, int frames_per_item,
                   int items_per_clip, int pair_ok, MelBank mel, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* s_wil = smem;                                             // [kPieceTaps][kMaxPieces] mel weights, tap-major
  int* s_pk0 = reinterpret_cast<int*>(s_wil + kPieceTaps * kMaxPieces);      // [kMaxPieces] first FFT bin of a piece
  int* s_pn = s_pk0 + kMaxPieces;                                  // [kMaxPieces] taps in the piece
  int* s_first = s_pn + kMaxPieces;                                // [kMaxMels + 4] first piece of every filter
  float* s_warp = reinterpret_cast<float*>(s_first + kMaxMels + 4);          // per warp: T, P, part
  __shared__ int s_npieces;
  __shared__ int s_round[kMaxPieces / 32];
  const int tid = threadIdx.x, warp = tid >> 5, L = tid & 31;
  float2* T = reinterpret_cast<float2*>(s_warp + warp * (2 * 16 * kTS + kPLen + kMaxPieces));
  float* P = reinterpret_cast<float*>(T + 16 * kTS);
  float* part = P + kPLen;

  // ---- mel projection schedule (see kPieceTaps), once per CTA
  if (!kPower) {
    if (tid == 0) {
      int np = 0;
      for (int m = 0; m < mel.n_mels; ++m) {
        s_first[m] = np;
        const int lo = mel.lo[m], o0 = mel.off[m], n = mel.off[m + 1] - o0;
        for (int c = 0; c < n; c += kPieceTaps) {
          s_pk0[np] = lo + c;
          s_pn[np] = min(kPieceTaps, n - c);
          ++np;
        }
      }
      s_first[mel.n_mels] = np;
      s_npieces = np;
      // round q = pieces [32 q, 32 q + 32): one per lane.  Its trip count is the longest piece in it; shorter pieces
      // run on zero weights (s_wil is zero beyond a piece's taps), so the loop bound is uniform and unrollable.
      for (int q = 0; q < kMaxPieces / 32; ++q) {
        int mx = 0;
        for (int pp = 32 * q; pp < min(np, 32 * q + 32); ++pp) mx = max(mx, s_pn[pp]);
        s_round[q] = mx;
      }
      for (int pp = np; pp < kMaxPieces; ++pp) { s_pk0[pp] = 0; s_pn[pp] = 0; }
    }
    for (int i = tid; i < kPieceTaps * kMaxPieces; i += kThreadsW) s_wil[i] = 0.f;
    for (int i = tid; i < kWarpsW * (kPLen - kBins); i += kThreadsW)          // the tail a zero-weight tap may read
      s_warp[(i / (kPLen - kBins)) * (2 * 16 * kTS + kPLen + kMaxPieces) + 2 * 16 * kTS + kBins + i % (kPLen - kBins)] = 0.f;
    __syncthreads();
    for (int m = tid; m < mel.n_mels; m += kThreadsW) {
      const int o0 = mel.off[m], n = mel.off[m + 1] - o0, p0 = s_first[m];
      for (int i = 0; i < n; ++i) s_wil[(i % kPieceTaps) * kMaxPieces + p0 + i / kPieceTaps] = mel.w[o0 + i];
    }
    __syncthreads();
  }
  const int n_pieces = kPower ? 0 : s_npieces;

  // ---- per-lane constants
  const int r_ = L & 15, h_ = L >> 4;
#if SED_LOGMEL_SMEM_WIN
  float2* s_win = reinterpret_cast<float2*>(s_warp + kWarpsW * (2 * 16 * kTS + kPLen + kMaxPieces));   // [512] window pairs
  for (int i = tid; i < kHalf; i += kThreadsW)
    s_win[i] = make_float2(0.25f - 0.25f * cospif(2.0f * (float)(2 * i) / (float)kNfft),
                           0.25f - 0.25f * cospif(2.0f * (float)(2 * i + 1) / (float)kNfft));
  __syncthreads();
#else
  float2 win[16];           // 0.5 * periodic Hann at samples 2 (L + 32 m), 2 (L + 32 m) + 1 (the 0.5 is the 1/2 of the real-FFT split)
#endif
  float2 w3[8];             // W32^(i + 8 h)
  float2 w5[15];            // W512^(r L), r = 1..15
  float2 w7[8];             // -i W1024^(L + 32 kp)
#if !SED_LOGMEL_SMEM_WIN
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int n2 = 2 * (L + 32 * m);
    win[m] = make_float2(0.25f - 0.25f * cospif(2.0f * (float)n2 / (float)kNfft),
                         0.25f - 0.25f * cospif(2.0f * (float)(n2 + 1) / (float)kNfft));
  }
#endif
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    w3[i] = twiddle(i + 8 * h_, 32);
    const float2 wk = twiddle(L + 32 * i, kNfft);
    w7[i] = make_float2(wk.y, -wk.x);                              // -i (x + i y) = y - i x
  }
#pragma unroll
  for (int r = 1; r < 16; ++r) w5[r - 1] = twiddle((r * L) & 511, 512);

  const long long total_items = (long long)n_clips * items_per_clip;
  for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int clip = (int)(item / items_per_clip);
    const int frame0 = (int)(item % items_per_clip) * frames_per_item;
    const int frames_here = min(frames_per_item, n_frames - frame0);
    const InT* __restrict__ src = wave + (long long)clip * n_samples;
    for (int f = warp; f < frames_here; f += kWarpsW) {
      const int frame = frame0 + f;
      const long long i0 = (long long)frame * hop - kHalf;         // sample index of the frame's first tap
      float2 v[16];
      // ---- load + window
      if (pair_ok && i0 >= 0 && i0 + kNfft <= n_samples) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const float2 xs = load_pair<InT>(src, i0 + 2 * (L + 32 * m));
#if SED_LOGMEL_SMEM_WIN
          const float2 wv = s_win[L + 32 * m];
#else
          const float2 wv = win[m];
#endif
          v[m] = make_float2(xs.x * wv.x, xs.y * wv.y);
        }
      } else {                                                      // clip edges: reflect padding, element by element
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          float xs[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            long long i = i0 + 2 * (L + 32 * m) + e;
            if (i < 0) i = -i;
            if (i >= n_samples) i = 2LL * (n_samples - 1) - i;
            xs[e] = (i >= 0 && i < n_samples) ? Sample<InT>::cvt(src[i]) : 0.f;
          }
#if SED_LOGMEL_SMEM_WIN
          const float2 wv = s_win[L + 32 * m];
#else
          const float2 wv = win[m];
#endif
          v[m] = make_float2(xs[0] * wv.x, xs[1] * wv.y);
        }
      }
      // ---- DFT16 over m:  E_h[q]
      dft16(v);
      // ---- 32-point DFT over j' = h + 2 m for row r: lanes L and L ^ 16 exchange half of their E
      float2 zl[8], zh[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 snd = sel2(h_ != 0, v[i], v[i + 8]);
        const float2 rcv = shfl2_xor(snd, 16);
        const float2 a = sel2(h_ != 0, rcv, v[i]);
        const float2 b = sel2(h_ != 0, v[i + 8], rcv);
        const float2 t = cmul(b, w3[i]);
        zl[i] = cadd(a, t);                                         // kj = i + 8 h
        zh[i] = csub(a, t);                                         // kj = i + 8 h + 16
      }
      // ---- transpose: lane (r, h) -> T[r][kj];  lane kj <- T[r][kj], r = 0..15
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        T[r_ * kTS + i + 8 * h_] = zl[i];
        T[r_ * kTS + i + 8 * h_ + 16] = zh[i];
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = T[r * kTS + L];
      __syncwarp();
#pragma unroll
      for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], w5[r - 1]);
      // ---- DFT16 over r:  v[kr'] = Z[L + 32 kr']  (half scale)
      dft16(v);
      // ---- real-FFT split + power:  bins k = L + 32 kp (kp < 8) and 512 - k
      const int src_lane = (32 - L) & 31;
#pragma unroll
      for (int kp = 0; kp < 8; ++kp) {
        const float2 snd = sel2(L == 0, v[(16 - kp) & 15], v[15 - kp]);
        float2 zm = shfl2(snd, src_lane);                           // Z[512 - k]
        zm.y = -zm.y;
        const float2 e = cadd(v[kp], zm), d = csub(v[kp], zm);
        const float2 t = cmul(d, w7[kp]);
        const float2 xp = cadd(e, t), xm = csub(e, t);
        P[L + 32 * kp] = xp.x * xp.x + xp.y * xp.y;
        P[kHalf - (L + 32 * kp)] = xm.x * xm.x + xm.y * xm.y;
      }
      if (L == 0) P[256] = 4.0f * (v[8].x * v[8].x + v[8].y * v[8].y);   // self-paired bin: |X[256]|^2 = |Z[256]|^2
      __syncwarp();
      const long long gframe = (long long)clip * n_frames + frame;
      if (kPower) {
        float* dst = out + gframe * kBins;
        for (int k = L; k < kBins; k += 32) dst[k] = P[k];
      } else {
#pragma unroll
        for (int q = 0; q < kMaxPieces / 32; ++q) {
          if (32 * q < n_pieces) {
            const int p = 32 * q + L, k0 = s_pk0[p], cnt = s_round[q];
            float acc = 0.f;
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) acc = fmaf(s_wil[i * kMaxPieces + p], P[k0 + i], acc);
            part[p] = acc;
          }
        }
        __syncwarp();
        float* dst = out + gframe * mel.n_mels;
        for (int m = L; m < mel.n_mels; m += 32) {
          float acc = 0.f;
          for (int p = s_first[m]; p < s_first[m + 1]; ++p) acc += part[p];
          dst[m] = 10.0f * log10f(fmaxf(acc, mel.amin)) - mel.db_offset;
        }
      }
      __syncwarp();
    }
  }
}

template <typename InT, bool kPower>
int launch_logmel(const InT* wave, int n_clips, int n_samples, int hop, MelBank mel, float* out,
                  cudaStream_t stream, const char* name) {
  SED_REQUIRE(n_clips >= 0 && n_samples > kHalf, "%s: need n_samples > %d for reflect padding (got %d)",
              name, kHalf, n_samples);
  SED_REQUIRE(hop > 0 && hop % 2 == 0 && hop <= 2048, "%s: hop must be even and in (0, 2048] (got %d)", name, hop);
  SED_REQUIRE(aligned(out, 4) && aligned(wave, sizeof(InT)), "%s: misaligned pointer", name);
  if (n_clips == 0) return 0;
  SED_REQUIRE(wave && out, "%s: null pointer", name);
  SED_REQUIRE(kPower || mel.n_mels <= kMaxMels, "%s: at most %d mel filters (got %d)", name, kMaxMels, mel.n_mels);
  const int n_frames = n_samples / hop + 1;
  // consecutive frames of an item are walked by the four warps of one CTA (L1 serves their overlap); small inputs get
  // small items so that every SM has work
  const int per_sm = SED_LOGMEL_CTAS;
  int fpi = kFramesPerChunk;
  while (fpi > kWarpsW && (long long)n_clips * ceil_div(n_frames, fpi) < 2LL * sm_count() * per_sm) fpi >>= 1;
  const int items_per_clip = ceil_div(n_frames, fpi);
  const size_t smem = sizeof(float) * (size_t)(kPieceTaps * kMaxPieces + 2 * kMaxPieces + kMaxMels + 4 +
                                               kWarpsW * (2 * 16 * kTS + kPLen + kMaxPieces) +
                                               (SED_LOGMEL_SMEM_WIN ? 2 * kHalf : 0));
  // 8-byte (fp32) / 4-byte (int16) pair loads need an even clip length and an aligned base
  const int pair_ok = (n_samples % 2 == 0 && aligned(wave, 2 * sizeof(InT))) ? 1 : 0;
  auto kern = logmel_warp_kernel<InT, kPower>;
  SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)n_clips * items_per_clip;
  const int grid = (int)min(items, (long long)sm_count() * per_sm);
  kern<<<grid, kThreadsW, smem, stream>>>(wave, n_clips, n_samples, hop, n_frames, fpi, items_per_clip, pair_ok, mel, out);
  SED_LAUNCH_CHECK(name);
  return 0;
}

// standalone mel projection + dB (unfused seam A, LogmelFilterBank.forward on its own)
__global__ void mel_db_kernel(const float* __restrict__ power, long long rows, int n_bins, MelBank mel,
                              int is_log, float* __restrict__ out) {
  extern __shared__ float s_p[];
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    __syncthreads();
    for (int k = threadIdx.x; k < n_bins; k += blockDim.x) s_p[k] = power[row * n_bins + k];
    __syncthreads();
    for (int m = threadIdx.x; m < mel.n_mels; m += blockDim.x) {
      const int lo = mel.lo[m], o0 = mel.off[m], o1 = mel.off[m + 1];
      float acc = 0.f;
      for (int t = o0; t < o1; ++t) acc = fmaf(mel.w[t], s_p[lo + (t - o0)], acc);
      out[row * mel.n_mels + m] = is_log ? 10.0f * log10f(fmaxf(acc, mel.amin)) - mel.db_offset : acc;
    }
  }
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_logmel_f32(const float* wave, int n_clips, int n_samples, int hop, const float* mel_w,
                   const int* mel_lo, const int* mel_off, int n_mels, float amin, float db_offset,
                   float* out, sed_stream_t stream) {
  SED_REQUIRE(mel_w && mel_lo && mel_off && n_mels > 0, "sed_logmel_f32: bad mel bank");
  MelBank mel{mel_w, mel_lo, mel_off, n_mels, amin, db_offset};
  return launch_logmel<float, false>(wave, n_clips, n_samples, hop, mel, out, (cudaStream_t)stream,
                                     "sed_logmel_f32");
}

int sed_logmel_i16(const int16_t* pcm, int n_clips, int n_samples, int hop, const float* mel_w,
                   const int* mel_lo, const int* mel_off, int n_mels, float amin, float db_offset,
                   float* out, sed_stream_t stream) {
  SED_REQUIRE(mel_w && mel_lo && mel_off && n_mels > 0, "sed_logmel_i16: bad mel bank");
  MelBank mel{mel_w, mel_lo, mel_off, n_mels, amin, db_offset};
  return launch_logmel<int16_t, false>(pcm, n_clips, n_samples, hop, mel, out, (cudaStream_t)stream,
                                       "sed_logmel_i16");
}

int sed_stft_power_f32(const float* wave, int n_clips, int n_samples, int hop, float* out_power,
                       sed_stream_t stream) {
  MelBank mel{nullptr, nullptr, nullptr, 0, 0.f, 0.f};
  return launch_logmel<float, true>(wave, n_clips, n_samples, hop, mel, out_power, (cudaStream_t)stream,
                                    "sed_stft_power_f32");
}

int sed_mel_db_f32(const float* power, long long rows, int n_bins, const float* mel_w, const int* mel_lo,
                   const int* mel_off, int n_mels, float amin, float db_offset, int is_log, float* out,
                   sed_stream_t stream) {
  SED_REQUIRE(power && out && mel_w && mel_lo && mel_off, "sed_mel_db_f32: null pointer");
  SED_REQUIRE(n_bins > 0 && n_bins <= 8192 && n_mels > 0, "sed_mel_db_f32: bad sizes");
  if (rows == 0) return 0;
  MelBank mel{mel_w, mel_lo, mel_off, n_mels, amin, db_offset};
  const int grid = (int)min(rows, (long long)sm_count() * 16);
  mel_db_kernel<<<grid, 128, n_bins * sizeof(float), (cudaStream_t)stream>>>(power, rows, n_bins, mel, is_log, out);
  SED_LAUNCH_CHECK("sed_mel_db_f32");
  return 0;
}

}  // extern "C"
