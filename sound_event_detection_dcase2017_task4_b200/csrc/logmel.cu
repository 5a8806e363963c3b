// logmel.cu -- fused STFT(1024, Hann, reflect) -> |.|^2 -> sparse mel -> 10*log10 front-end.
//
// Replaces torchlibrosa.stft.Spectrogram + LogmelFilterBank as called at
// /root/reference/pytorch/models.py:199-200 (ctor contracts :166-173).  The reference evaluates
// the windowed DFT as two dense fp32 conv1d (2.1 GFLOP per 10 s clip) and materialises the
// (T,513) power spectrogram; here every frame is a shared-memory radix-8 real FFT and only the
// (T,64) log-mel leaves the SM.
//
// Work decomposition: a CTA (256 threads) owns FPC consecutive frames of one clip.  The frames
// overlap 1024/hop = 3.2x, so the CTA loads the contiguous sample span once (coalesced float4 /
// int4, reflect padding folded into the gather) and runs 4 frames at a time, 64 threads each:
//   real 1024-FFT  =  512-point complex FFT of z[n] = x[2n] + i x[2n+1]  (8 x 8 x 8, three
//   radix-8 passes, one complex value per (thread, r)), then the even/odd split.
// Persistent grid: CTAs stride over (clip, chunk) work items so the twiddle / window prologue
// is paid once per CTA.
#include "common.cuh"

namespace sed {
namespace {

constexpr int kNfft = 1024;
constexpr int kBins = kNfft / 2 + 1;   // 513
constexpr int kHalf = kNfft / 2;       // 512 complex points
constexpr int kThreads = 256;
constexpr int kSlots = 4;              // frames in flight per CTA
constexpr int kFramesPerChunk = 16;
constexpr int kBuf = 8 * 72;           // one padded SoA plane of the 512-point work buffer
// Mel projection schedule (built once per CTA in shared memory): every mel filter is cut into pieces of at most
// kPieceTaps consecutive taps; piece p is owned by thread p % 64 of a frame slot.  The reference bank has 866 taps in
// filters of 3..47 taps: one thread per filter made the 47-tap thread the critical path and every weight a scattered
// 4-byte global load (ncu r01: 78 % excessive sectors, L1/TEX 80 % busy).  Pieces balance the slot (<= 22 taps per
// thread instead of 47) and the weights are staged tap-major / piece-minor so a warp reads them without conflicts.
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

// the 64 threads (2 warps) of one frame slot only exchange data among themselves: a named barrier per slot lets the
// four slots of a CTA drift apart instead of meeting at a CTA-wide barrier five times per frame
__device__ __forceinline__ void slot_barrier(int slot) {
  asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");
}

__device__ __forceinline__ int zaddr(int k) { return k + 4 * (k >> 5); }   // padded Z layout (<576)

template <typename T> struct Sample;
template <> struct Sample<float> {
  static __device__ __forceinline__ float cvt(float v) { return v; }
};
template <> struct Sample<int16_t> {
  // utils/utilities.py:66-67: (x / 32767.) in float64 then cast to float32.  fp32 division of the
  // exactly-representable integer is the correctly rounded quotient; the double-then-float path
  // agrees with it for every int16 value (tests/test_host_logic.py checks all 65536).
  static __device__ __forceinline__ float cvt(int16_t v) { return __fdiv_rn((float)v, 32767.0f); }
};

struct MelBank {
  const float* w;
  const int* lo;
  const int* off;
  int n_mels;
  float amin, db_offset;
};

// kPower: write the 513-bin power spectrogram instead of the log-mel (unfused seam A).
template <typename InT, bool kPower>
__global__ void __launch_bounds__(kThreads, 3)
logmel_kernel(const InT* __restrict__ wave, int n_clips, int n_samples, int hop, int n_frames,
              int chunks_per_clip, MelBank mel, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* s_work = smem;                      // [kSlots][4][kBuf]  A.re A.im B.re B.im
  float* s_wil = s_work + kSlots * 4 * kBuf;   // [kPieceTaps][kMaxPieces] mel weights, tap-major (kPower: unused)
  float* s_part = s_wil + kPieceTaps * kMaxPieces;                 // [kSlots][kMaxPieces] per-piece partial sums
  int* s_pk0 = reinterpret_cast<int*>(s_part + kSlots * kMaxPieces);   // [kMaxPieces] first FFT bin of the piece
  int* s_pn = s_pk0 + kMaxPieces;              // [kMaxPieces] taps in the piece
  int* s_first = s_pn + kMaxPieces;            // [kMaxMels + 1] first piece of every mel filter
  float* s_span = reinterpret_cast<float*>(s_first + kMaxMels + 4);    // [(FPC-1)*hop + 1024], 16-byte aligned
  __shared__ int s_npieces;

  const int tid = threadIdx.x;
  const int slot = tid >> 6;
  const int j = tid & 63;
  float* A_re = s_work + slot * 4 * kBuf;
  float* A_im = A_re + kBuf;
  float* B_re = A_im + kBuf;
  float* B_im = B_re + kBuf;

  // ---- prologue: everything a thread needs for every frame lives in its registers -----------
  // (the window samples and split twiddles used to be re-read from shared memory for every frame: the kernel is
  //  shared-memory-bandwidth bound, so they are per-thread constants now)
  float2 tw1[8], tw2[8];                      // W512^(j q)  and  W64^(j0 p)
  float2 win[8];                              // periodic Hann at samples 2(j + 64 r), 2(j + 64 r) + 1
  float2 tws[4];                              // W1024^(j + 64 s), s = 0..3, for the even/odd split
  {
    const int j0 = j & 7;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      tw1[q] = twiddle((j * q) & 511, 512);
      tw2[q] = twiddle((j0 * q) & 63, 64);
      const int n2 = 2 * (j + 64 * q);
      win[q] = make_float2(0.5f - 0.5f * cospif(2.0f * (float)n2 / (float)kNfft),
                           0.5f - 0.5f * cospif(2.0f * (float)(n2 + 1) / (float)kNfft));
    }
#pragma unroll
    for (int sI = 0; sI < 4; ++sI) tws[sI] = twiddle(j + 64 * sI, kNfft);
  }

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
    }
    for (int i = tid; i < kPieceTaps * kMaxPieces; i += kThreads) s_wil[i] = 0.f;
    __syncthreads();
    // weights: tap t of filter m belongs to piece s_first[m] + (t - off[m]) / kPieceTaps, position (t - off[m]) % kPieceTaps
    for (int m = tid; m < mel.n_mels; m += kThreads) {
      const int o0 = mel.off[m], n = mel.off[m + 1] - o0, p0 = s_first[m];
      for (int i = 0; i < n; ++i) s_wil[(i % kPieceTaps) * kMaxPieces + p0 + i / kPieceTaps] = mel.w[o0 + i];
    }
    __syncthreads();
  }
  const int n_pieces = kPower ? 0 : s_npieces;

  const long long total_items = (long long)n_clips * chunks_per_clip;
  for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int clip = (int)(item / chunks_per_clip);
    const int chunk = (int)(item % chunks_per_clip);
    const int frame0 = chunk * kFramesPerChunk;
    const int frames_here = min(kFramesPerChunk, n_frames - frame0);
    const InT* __restrict__ src = wave + (long long)clip * n_samples;
    // ---- load the sample span [frame0*hop - 512, ...) with reflect padding ------------------
    const long long i0 = (long long)frame0 * hop - kHalf;          // original index of span[0]
    const int need = (frames_here - 1) * hop + kNfft;
    __syncthreads();                                               // previous item done with s_span
    const bool interior = (i0 >= 0) && (i0 + need <= n_samples);
    if (interior && sizeof(InT) == 4 && ((reinterpret_cast<uintptr_t>(src + i0) & 15) == 0)) {
      const float4* s4 = reinterpret_cast<const float4*>(src + i0);
      float4* d4 = reinterpret_cast<float4*>(s_span);
      for (int v = tid; v < (need >> 2); v += kThreads) d4[v] = __ldg(s4 + v);
      for (int s = (need & ~3) + tid; s < need; s += kThreads) s_span[s] = Sample<InT>::cvt(src[i0 + s]);
    } else if (interior && sizeof(InT) == 2 && ((reinterpret_cast<uintptr_t>(src + i0) & 15) == 0)) {
      const int4* s8 = reinterpret_cast<const int4*>(src + i0);
      for (int v = tid; v < (need >> 3); v += kThreads) {
        int4 raw = __ldg(s8 + v);
        const int16_t* p = reinterpret_cast<const int16_t*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) s_span[v * 8 + e] = Sample<int16_t>::cvt(p[e]);
      }
      for (int s = (need & ~7) + tid; s < need; s += kThreads) s_span[s] = Sample<InT>::cvt(src[i0 + s]);
    } else {
      for (int s = tid; s < need; s += kThreads) {
        long long i = i0 + s;
        if (i < 0) i = -i;
        if (i >= n_samples) i = 2LL * (n_samples - 1) - i;
        float v = 0.f;
        if (i >= 0 && i < n_samples) v = Sample<InT>::cvt(src[i]);
        s_span[s] = v;
      }
    }
    __syncthreads();

    for (int fbase = 0; fbase < frames_here; fbase += kSlots) {
      const int f = fbase + slot;                 // frame within the chunk
      const bool live = f < frames_here;
      float2 v[8];
      // ---- pass 1: DFT8 over r of z[j + 64 r], twiddle W512^(j q) -> A[q][j] -----------------
      if (live) {
        const float* x = s_span + f * hop;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int n2 = 2 * (j + 64 * r);             // 8-byte aligned: hop and n2 are even
          const float2 xs = *reinterpret_cast<const float2*>(x + n2);
          v[r] = make_float2(xs.x * win[r].x, xs.y * win[r].y);
        }
        dft8(v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float2 y = (q == 0) ? v[0] : cmul(v[q], tw1[q]);
          A_re[q * 72 + j] = y.x;
          A_im[q * 72 + j] = y.y;
        }
      }
      slot_barrier(slot);
      // ---- pass 2: thread (q, j0): DFT8 over j1 of A[q][j0 + 8 j1], twiddle W64^(j0 p) -> B --
      if (live) {
        const int q = j >> 3, j0 = j & 7;
#pragma unroll
        for (int j1 = 0; j1 < 8; ++j1) v[j1] = make_float2(A_re[q * 72 + j0 + 8 * j1], A_im[q * 72 + j0 + 8 * j1]);
        dft8(v);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          float2 u = (p == 0) ? v[0] : cmul(v[p], tw2[p]);
          B_re[q * 72 + j0 * 9 + p] = u.x;
          B_im[q * 72 + j0 * 9 + p] = u.y;
        }
      }
      slot_barrier(slot);
      // ---- pass 3: thread (q, p): DFT8 over j0 -> Z[q + 8 p + 64 s] -> A (padded linear) -----
      if (live) {
        const int q = j >> 3, p = j & 7;
#pragma unroll
        for (int j0 = 0; j0 < 8; ++j0) v[j0] = make_float2(B_re[q * 72 + j0 * 9 + p], B_im[q * 72 + j0 * 9 + p]);
        dft8(v);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int k = q + 8 * p + 64 * s;
          A_re[zaddr(k)] = v[s].x;
          A_im[zaddr(k)] = v[s].y;
        }
      }
      slot_barrier(slot);
      // ---- even/odd split -> power spectrum P[0..512] -> B_re ---------------------------------
      float* P = B_re;                              // 513 <= kBuf
      // Bins k and 512-k share their inputs: with E = (Z[k] + conj Z[512-k])/2, O = (Z[k] - conj Z[512-k])/(2i) and
      // t = W1024^k O,  X[k] = E + t  and  X[512-k] = conj(E - t).  Thread j takes k = j + 64 s, s = 0..3 (and their
      // partners 512-k; k = 0 pairs with the Nyquist bin), thread 0 also the self-paired k = 256.
      if (live) {
#pragma unroll
        for (int sI = 0; sI < 4; ++sI) {
          const int k = j + 64 * sI;
          const int km = (kHalf - k) & (kHalf - 1);
          const float2 zk = make_float2(A_re[zaddr(k)], A_im[zaddr(k)]);
          const float2 zm = make_float2(A_re[zaddr(km)], -A_im[zaddr(km)]);   // conj Z[512-k]
          const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
          const float2 d = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y - zm.y));
          const float2 t = cmul(tws[sI], mul_neg_i(d));                       // W^k (Zk - conj Zm) / (2i)
          const float2 xp = cadd(e, t), xm = csub(e, t);
          P[k] = xp.x * xp.x + xp.y * xp.y;
          P[kHalf - k] = xm.x * xm.x + xm.y * xm.y;
        }
        if (j == 0) {                               // k = 256: Z[256] pairs with itself, W1024^256 = -i
          const float2 z = make_float2(A_re[zaddr(256)], A_im[zaddr(256)]);
          P[256] = z.x * z.x + z.y * z.y;           // |Re z - i Im z|^2
        }
      }
      slot_barrier(slot);
      if (live) {
        const long long frame = (long long)clip * n_frames + frame0 + f;
        if (kPower) {
          float* dst = out + frame * kBins;
          for (int k = j; k < kBins; k += 64) dst[k] = P[k];
        }
      }
      if (!kPower) {
        // mel projection in two balanced steps: per-piece partial sums, then one thread per filter adds its pieces
        // in fixed order, converts to dB and stores (64 consecutive floats per frame).
        float* part = s_part + slot * kMaxPieces;
        if (live) {
          for (int p = j; p < n_pieces; p += 64) {
            const int k0 = s_pk0[p], n = s_pn[p];
            float acc = 0.f;
#pragma unroll 4
            for (int i = 0; i < n; ++i) acc = fmaf(s_wil[i * kMaxPieces + p], P[k0 + i], acc);
            part[p] = acc;
          }
        }
        slot_barrier(slot);
        if (live) {
          const long long frame = (long long)clip * n_frames + frame0 + f;
          float* dst = out + frame * mel.n_mels;
          for (int m = j; m < mel.n_mels; m += 64) {
            float acc = 0.f;
            for (int p = s_first[m]; p < s_first[m + 1]; ++p) acc += part[p];
            dst[m] = 10.0f * log10f(fmaxf(acc, mel.amin)) - mel.db_offset;
          }
        }
      }
      // next iteration's pass 1 only writes A; B (=P) is rewritten after the next barrier.
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
  const int n_frames = n_samples / hop + 1;
  const int chunks = ceil_div(n_frames, kFramesPerChunk);
  SED_REQUIRE(kPower || mel.n_mels <= kMaxMels, "%s: at most %d mel filters (got %d)", name, kMaxMels, mel.n_mels);
  const size_t smem = sizeof(float) * (size_t)(kSlots * 4 * kBuf + kPieceTaps * kMaxPieces + kSlots * kMaxPieces +
                                               2 * kMaxPieces + kMaxMels + 4 + (kFramesPerChunk - 1) * hop + kNfft);
  auto kern = logmel_kernel<InT, kPower>;
  static thread_local int configured_dev = -1;   // attribute is per (function, device)
  int dev = 0;
  SED_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_dev = dev;
  }
  SED_REQUIRE(smem <= 200 * 1024, "%s: hop %d needs %zu B of shared memory", name, hop, smem);
  const long long items = (long long)n_clips * chunks;
  const int per_sm = (int)max((size_t)1, min((size_t)3, (size_t)(220 * 1024) / smem));
  const int grid = (int)min(items, (long long)sm_count() * per_sm);
  kern<<<grid, kThreads, smem, stream>>>(wave, n_clips, n_samples, hop, n_frames, chunks, mel, out);
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
