// bnpool.cu -- training-mode BatchNorm2d + ReLU + average pooling, forward and backward, on
// NHWC bf16 activations (HBM-bound warp/CTA-reduction kernels).
//
// Replaces, for each of the 8 conv layers, `F.relu_(self.bnX(self.convX(x)))` followed by
// `F.avg_pool2d` (/root/reference/pytorch/models.py:102-113) and, for block 4, also the
// `torch.mean(x, dim=3)` at models.py:218/303/386/473/564/741/835; plus their autograd backward.
//
// Forward of one layer:
//   conv kernel epilogue   -> per-CTA (sum, sumsq) partials of the fp32 conv output
//   sed_bn_finalize        -> mean / biased var in fp64, scale = gamma*invstd, shift = beta - mean*scale,
//                             running stats update (momentum, unbiased var), num_batches_tracked += 1
//   sed_bn_relu_pool_fwd   -> out = avgpool_{ph x pw}( relu(y*scale + shift) )   (one read of y)
// Backward of one layer (two passes over y, no saved activation other than y itself):
//   sed_bn_relu_pool_bwd_reduce -> partials of  sum(g), sum(g*xhat),  g = unpool(dA)/(ph*pw) * [bn(y) > 0]
//   sed_bn_bwd_finalize         -> dgamma, dbeta and the three per-channel coefficients
//   sed_bn_relu_pool_bwd_apply  -> dY = gamma*invstd * (g - mean(g) - xhat*mean(g*xhat))      (bf16)
#include "common.cuh"
#include <string.h>

namespace sed {
namespace {

constexpr int kVec = 8;   // channels per thread: one 16-byte bf16 vector

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8f(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ---------------------------------------------------------------- finalize (statistics -> affine)
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int P, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_invstd) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, ss;
  colsum2_block(partial, partial + C, P, 2LL * C, c, c < C, s, ss);
  if (threadIdx.y != 0) return;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (c >= C) return;
  const double mean = s / count;
  double var = ss / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - (float)mean * g * invstd;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: affine from the running statistics
__global__ void bn_eval_affine_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      int C, float* __restrict__ scale, float* __restrict__ shift,
                                      float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.0f / sqrtf(running_var[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - running_mean[c] * g * invstd;
  if (save_mean) save_mean[c] = running_mean[c];
  if (save_invstd) save_invstd[c] = invstd;
}

// ---------------------------------------------------------------- forward
template <bool kOutF32>
__global__ void bn_relu_pool_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                        const float* __restrict__ shift, int B, int H, int W, int C, int ph, int pw,
                                        void* __restrict__ out_) {
  const int Ho = H / ph, Wo = W / pw, CV = C / kVec;
  const long long total = (long long)B * Ho * Wo * CV;
  const float inv = 1.0f / (float)(ph * pw);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long r = i / CV;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float sc[8], sh[8], acc[8];
    load8f(scale + cv * kVec, sc);
    load8f(shift + cv * kVec, sh);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int dh = 0; dh < ph; ++dh)
      for (int dw = 0; dw < pw; ++dw) {
        float v[8];
        load8(y + ((((long long)b * H + ho * ph + dh) * W + wo * pw + dw) * C + cv * kVec), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f);
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv;
    const long long o = (((long long)b * Ho + ho) * Wo + wo) * C + cv * kVec;
    if (kOutF32) store8f(reinterpret_cast<float*>(out_) + o, acc);
    else store8(reinterpret_cast<__nv_bfloat16*>(out_) + o, acc);
  }
}

// ---------------------------------------------------------------- backward pass 1: reductions
// block = 256 threads = CV channel-vectors x (256/CV) pixel lanes; each block owns a contiguous
// range of input pixels (b,h,w) and writes one [2][C] partial.
template <bool kGradF32>
__global__ void bn_relu_pool_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                               int B, int H, int W, int C, int ph, int pw,
                                               float* __restrict__ partial) {
  extern __shared__ float s_red[];                       // [lanes][2*C]
  const int Ho = H / ph, Wo = W / pw, CV = C / kVec;
  const int lanes = blockDim.x / CV;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const long long npix = (long long)B * H * W;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  const float inv = 1.0f / (float)(ph * pw);
  float sc[8], sh[8], sg[8], sgx[8];
  load8f(scale + cv * kVec, sc);
  load8f(shift + cv * kVec, sh);
#pragma unroll
  for (int k = 0; k < 8; ++k) sg[k] = sgx[k] = 0.f;
  if (pl < lanes) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      const int w = (int)(p % W);
      const long long r = p / W;
      const int h = (int)(r % H);
      const int b = (int)(r / H);
      const int ho = h / ph, wo = w / pw;
      if (ho >= Ho || wo >= Wo) continue;                // floor-mode tail: no gradient
      float v[8], g[8];
      load8(y + p * C + cv * kVec, v);
      const long long o = (((long long)b * Ho + ho) * Wo + wo) * C + cv * kVec;
      if (kGradF32) load8f(reinterpret_cast<const float*>(dA_) + o, g);
      else load8(reinterpret_cast<const __nv_bfloat16*>(dA_) + o, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float act = fmaf(v[k], sc[k], sh[k]);
        const float gk = act > 0.f ? g[k] * inv : 0.f;
        sg[k] += gk;
        sgx[k] = fmaf(gk, v[k], sgx[k]);               // raw y: sed_bn_bwd_finalize converts with the batch mean
      }
    }
  }
  if (pl < lanes) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s_red[pl * 2 * C + cv * kVec + k] = sg[k];
      s_red[pl * 2 * C + C + cv * kVec + k] = sgx[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * 2 * C + i];
    partial[(long long)blockIdx.x * 2 * C + i] = a;
  }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int P, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ invstd,
                                       const float* __restrict__ mean_for_gy,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                       int frozen_stats,
                                       float* __restrict__ coef /* [3][C]: gamma*invstd, mean(g), mean(g*xhat) */) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, sx;
  colsum2_block(partial, partial + C, P, 2LL * C, c, c < C, s, sx);
  if (threadIdx.y != 0 || c >= C) return;
  if (mean_for_gy) sx = (double)invstd[c] * (sx - (double)mean_for_gy[c] * s);    // partials held sum(g*y)
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)sx : (float)sx;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
  if (coef) {
    coef[c] = (gamma ? gamma[c] : 1.f) * invstd[c];
    // frozen (eval-mode) statistics are constants of the graph: dY = gamma * invstd * g, no batch terms
    coef[C + c] = frozen_stats ? 0.f : (float)(s / count);
    coef[2 * C + c] = frozen_stats ? 0.f : (float)(sx / count);
  }
}

// ---------------------------------------------------------------- backward pass 2: dY
template <bool kGradF32>
__global__ void bn_relu_pool_bwd_apply_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                                              const float* __restrict__ scale, const float* __restrict__ shift,
                                              const float* __restrict__ mean, const float* __restrict__ invstd,
                                              const float* __restrict__ coef, int B, int H, int W, int C, int ph,
                                              int pw, __nv_bfloat16* __restrict__ dy) {
  const int Ho = H / ph, Wo = W / pw, CV = C / kVec;
  const long long total = (long long)B * H * W * CV;
  const float inv = 1.0f / (float)(ph * pw);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const long long p = i / CV;
    const int w = (int)(p % W);
    const long long r = p / W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    const int ho = h / ph, wo = w / pw;
    float sc[8], sh[8], mu[8], is[8], c1[8], c2[8], c3[8], v[8], g[8], o[8];
    load8f(scale + cv * kVec, sc);
    load8f(shift + cv * kVec, sh);
    load8f(mean + cv * kVec, mu);
    load8f(invstd + cv * kVec, is);
    load8f(coef + cv * kVec, c1);
    load8f(coef + C + cv * kVec, c2);
    load8f(coef + 2 * C + cv * kVec, c3);
    load8(y + p * C + cv * kVec, v);
    const bool inside = ho < Ho && wo < Wo;
    if (inside) {
      const long long oi = (((long long)b * Ho + ho) * Wo + wo) * C + cv * kVec;
      if (kGradF32) load8f(reinterpret_cast<const float*>(dA_) + oi, g);
      else load8(reinterpret_cast<const __nv_bfloat16*>(dA_) + oi, g);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float act = fmaf(v[k], sc[k], sh[k]);
      const float gk = (inside && act > 0.f) ? g[k] * inv : 0.f;
      const float xhat = (v[k] - mu[k]) * is[k];
      o[k] = c1[k] * (gk - c2[k] - xhat * c3[k]);
    }
    store8(dy + p * C + cv * kVec, o);
  }
}

// ================================================================ row-structured fast paths
// Used when 256 % (C/8) == 0 and the pooling is 1x1 or 2x2 with even W: a CTA walks image rows,
// thread t always owns channel octet t % (C/8), so every per-channel constant lives in registers
// for the whole kernel and the inner loop has no integer division.  All accesses are 16-byte
// vectors, several independent loads in flight per thread.
constexpr int kRowThreads = 256;

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]) {
  float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// out row (b, ho) <- relu(bn(y rows ho*P .. ho*P+P-1)), P = 1 or 2 (square pool), bf16 out
template <int kPool>
__global__ void __launch_bounds__(kRowThreads, 4)
bn_relu_pool_fwd_rows_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                             const float* __restrict__ shift, int B, int H, int W, int C,
                             __nv_bfloat16* __restrict__ out) {
  const int CV = C >> 3, Ho = H / kPool, Wo = W / kPool;
  const int cv = threadIdx.x % CV, cv_shift = __ffs(CV) - 1;
  float sc[8], sh[8];
  load8f(scale + cv * 8, sc);
  load8f(shift + cv * 8, sh);
  const int out_vecs = Wo * CV;                    // vectors per output row
  const long long in_row = (long long)W * C;       // elements per input row
  const int rows = B * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / Ho, ho = row - b * Ho;
    const __nv_bfloat16* src = y + ((long long)b * H + (long long)ho * kPool) * in_row;
    __nv_bfloat16* dst = out + (long long)row * Wo * C;
#pragma unroll 2
    for (int v = threadIdx.x; v < out_vecs; v += kRowThreads) {
      const int wo = v >> cv_shift;                  // CV divides 256 => power of two
      uint4 raw[kPool * kPool];
#pragma unroll
      for (int dh = 0; dh < kPool; ++dh)
#pragma unroll
        for (int dw = 0; dw < kPool; ++dw)
          raw[dh * kPool + dw] = ldg16(src + dh * in_row + ((long long)(wo * kPool + dw) * CV + cv) * 8);
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
      for (int i = 0; i < kPool * kPool; ++i) {
        float x[8];
        unpack8(raw[i], x);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += fmaxf(fmaf(x[k], sc[k], sh[k]), 0.f);
      }
      if (kPool == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= 0.25f;
      }
      store8(dst + (long long)v * 8, acc);
    }
  }
}

// ---- un-pooled (1x1) layers as flat streams ----------------------------------------------------------------------
// Without pooling y, dA, dY and the activation are the same contiguous (pixels x C) array, so the kernels walk it
// as ONE stream of 16-byte vectors: thread t's vectors are t, t + stride, ... and, because C/8 divides the block
// size and the stride, they all belong to the same channel octet (per-channel constants stay in registers).  The
// row-structured versions had at most two loop iterations per image row (W*C/8 = 512 vectors per 256 threads) and
// the load pipeline drained at every row change; here 4 independent iterations are always in flight.
__global__ void __launch_bounds__(kRowThreads, 4)
bn_relu_fwd_flat_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                        const float* __restrict__ shift, long long nvec, int C, __nv_bfloat16* __restrict__ out) {
  const int CV = C >> 3, cv = threadIdx.x % CV;
  float sc[8], sh[8];
  load8f(scale + cv * 8, sc);
  load8f(shift + cv * 8, sh);
  const long long stride = (long long)gridDim.x * kRowThreads;
#pragma unroll 4
  for (long long v = (long long)blockIdx.x * kRowThreads + threadIdx.x; v < nvec; v += stride) {
    float x[8], o[8];
    unpack8(ldg16(y + v * 8), x);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaxf(fmaf(x[k], sc[k], sh[k]), 0.f);
    store8(out + v * 8, o);
  }
}

// ================================================================ backward fast paths (round 2)
// Co-residency-friendly, dynamically scheduled versions.  The weight-gradient kernel of layer l (tensor-pipe bound, one
// 192-thread CTA per SM holding ~210 KB of shared memory) only depends on dY_l, while the BatchNorm backward of layer
// l-1 (HBM bound) is the next link of the critical chain; engine.trunk_backward runs the two on different streams.
// For them to really share an SM these kernels
//   * stay at <= 64 registers / thread (a thread owns FOUR channels = 8-byte vectors, so the per-channel constants of
//     the apply pass fit), i.e. 3 CTAs fit beside a weight-gradient CTA, 4 when alone;
//   * use 4 KB of shared memory (the reduction scratch), which fits in what the weight-gradient kernel leaves free;
//   * take their work from an atomic ticket counter instead of assuming a full resident wave: a *virtual worker* w owns
//     a fixed contiguous range of vectors / pooled rows whatever CTA executes it, so the per-worker partial sums -- and
//     with them the fixed-order fp64 finalize -- stay bit-reproducible under any co-scheduling.  Worker sizes are
//     two-tier (the last quarter of the tensor is cut into 4x smaller workers) to keep the tail short.
// The reduction pass produces sum(g) and sum(g*y) (raw y); sed_bn_bwd_finalize converts with the batch mean.
constexpr int kDynThreads = 256;
#ifndef SED_BN_WIN_UNROLL
#define SED_BN_WIN_UNROLL 4      // measured: 1.55 -> 1.35 ms for the (1001 x 64, C = 64) layer vs 2
#endif
#ifndef SED_BN_FLAT_UNROLL
#define SED_BN_FLAT_UNROLL 4
#endif
constexpr int kFlatUnroll = SED_BN_FLAT_UNROLL;
constexpr int kWinUnroll = SED_BN_WIN_UNROLL;   // independent 2x2 windows in flight per thread (window kernels)

struct DynPlan {
  long long n_units;      // flat: blocks of 1024 four-channel vectors; win2: pooled rows
  long long big, small;   // units per worker in the two tiers
  int v_big, V;
};

__device__ __forceinline__ void dyn_range(const DynPlan& q, int w, long long& u0, long long& u1) {
  if (w < q.v_big) { u0 = (long long)w * q.big; u1 = u0 + q.big; }
  else { u0 = (long long)q.v_big * q.big + (long long)(w - q.v_big) * q.small; u1 = u0 + q.small; }
  if (u1 > q.n_units) u1 = q.n_units;
  if (u0 > u1) u0 = u1;
}
// ticket of the next virtual worker (block-uniform).  sched[0] = next ticket, sched[1] = CTAs that ran dry; both are
// zero before the launch and the last CTA to run dry re-zeroes them, so consecutive launches can share the words.
__device__ __forceinline__ int dyn_next(int* sched) {
  __shared__ int s_ticket;
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&sched[0], 1);
  __syncthreads();
  return s_ticket;
}
__device__ __forceinline__ void dyn_retire(int* sched) {
  if (threadIdx.x == 0) {
    const int d = atomicAdd(&sched[1], 1);
    if (d == (int)gridDim.x - 1) { atomicExch(&sched[0], 0); atomicExch(&sched[1], 0); }
  }
}

// 8-byte streaming load.  y / dA are read exactly once per pass: keep them out of L1 (no_allocate) so that the lines do
// not compete with the shared-memory traffic of a co-resident tensor-core CTA (the carve-out is one SRAM).
#ifndef SED_BN_LD_MODE
#define SED_BN_LD_MODE 1
#endif
__device__ __forceinline__ uint2 ldg8(const void* p) {
#if SED_BN_LD_MODE == 1
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const uint2*>(p));
#endif
}
__device__ __forceinline__ void unpack4(const uint2& raw, float (&v)[4]) {
  const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ void load4f(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <bool kGradF32>
__device__ __forceinline__ void load_grad4(const void* dA, long long vec, float (&g)[4]) {
  if (kGradF32) load4f(reinterpret_cast<const float*>(dA) + vec * 4, g);
  else unpack4(ldg8(reinterpret_cast<const __nv_bfloat16*>(dA) + vec * 4), g);
}

// per-worker partial row: row[0][c] = sum g, row[1][c] = sum g*y.  Thread (pl, cv) holds 4 channels of lane pl; the
// `lanes` = 256 / (C/4) lanes are summed in fixed order through a 4 KB scratch, one statistic at a time.
__device__ __forceinline__ void dyn_write_partial(float* s_red, const float (&sg)[4], const float (&sgy)[4], int C, int cv,
                                                  int pl, int lanes, float* __restrict__ row) {
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass) __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) s_red[pl * C + cv * 4 + k] = pass ? sgy[k] : sg[k];
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += kDynThreads) {
      float a = 0.f;
      for (int l = 0; l < lanes; ++l) a += s_red[l * C + i];
      row[pass * C + i] = a;
    }
  }
}
// Second level, still inside the reduce kernel: workers are grouped by kGroup consecutive ids; the CTA that completes
// the LAST worker of a group (whichever it is) adds the group's rows in worker order into one group row.  The sum is a
// fixed-order sum of fixed operands, so it does not depend on the arrival order; the finalize kernel then reads
// V / kGroup rows instead of V (with ~4000 workers it had become the longest link of the chain between two passes).
// Layout of the workspace: group rows [0, G), worker rows [G, G + V).  gctr: G zeroed counters, re-zeroed here.
constexpr int kGroup = 16;
constexpr int kMaxGroups = 2046;      // sched workspace = 2 + kMaxGroups int32 words
__device__ __forceinline__ void dyn_group_reduce(float* __restrict__ ws, int w, int V, int G, int C, int* gctr) {
  __shared__ int s_last;
  const int g = w / kGroup;
  const int first = g * kGroup, count = min(kGroup, V - first);
  __threadfence();                                   // this CTA's row is visible device-wide ...
  __syncthreads();                                   // ... for every thread of it, before the arrival is counted
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(&gctr[g], 1);
    s_last = arrived == count - 1;
    if (s_last) atomicExch(&gctr[g], 0);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* rows = ws + ((long long)G + first) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += kDynThreads) {
    float a = 0.f;
    for (int r = 0; r < count; ++r) a += __ldcg(rows + (long long)r * 2 * C + i);
    ws[(long long)g * 2 * C + i] = a;
  }
}

// ---- un-pooled layers (and block 4's whole-row average): y, dA, dY are one flat stream of 8-byte vectors
template <bool kGradF32>
__global__ void __launch_bounds__(kDynThreads, 4)
bn_bwd_reduce_flat4_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                           const float* __restrict__ scale, const float* __restrict__ shift, long long nvec, int C,
                           int row_shift, float row_scale, DynPlan q, int G, float* __restrict__ partial, int* sched) {
  __shared__ float s_red[kDynThreads * 4];
  const int CV = C >> 2, cv = threadIdx.x & (CV - 1), pl = threadIdx.x / CV, lanes = kDynThreads / CV;
  float sc[4], sh[4];
  load4f(scale + cv * 4, sc);
  load4f(shift + cv * 4, sh);
  for (;;) {
    const int w = dyn_next(sched);
    if (w >= q.V) break;
    long long u0, u1;
    dyn_range(q, w, u0, u1);
    const long long v1 = min(nvec, u1 * 1024);
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll kFlatUnroll
    for (long long v = u0 * 1024 + threadIdx.x; v < v1; v += kDynThreads) {
      const uint2 raw = ldg8(y + v * 4);
      const long long gv = row_shift >= 0 ? ((v >> row_shift) * CV + cv) : v;
      float g[4], x[4];
      load_grad4<kGradF32>(dA_, gv, g);
      unpack4(raw, x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = fmaf(x[k], sc[k], sh[k]) > 0.f ? g[k] * row_scale : 0.f;
        sg[k] += gk;
        sgy[k] = fmaf(gk, x[k], sgy[k]);
      }
    }
    dyn_write_partial(s_red, sg, sgy, C, cv, pl, lanes, partial + ((long long)G + w) * 2 * C);
    dyn_group_reduce(partial, w, q.V, G, C, sched + 2);
  }
  dyn_retire(sched);
}

template <bool kGradF32>
__global__ void __launch_bounds__(kDynThreads, 4)
bn_bwd_apply_flat4_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd,
                          const float* __restrict__ coef, long long nvec, int C, int row_shift, float row_scale,
                          DynPlan q, __nv_bfloat16* __restrict__ dy, int* sched) {
  const int CV = C >> 2, cv = threadIdx.x & (CV - 1);
  float sc[4], sh[4], cA[4], cB[4], cC[4];
  {
    float mu[4], is[4], c2[4], c3[4];
    load4f(scale + cv * 4, sc);
    load4f(shift + cv * 4, sh);
    load4f(mean + cv * 4, mu);
    load4f(invstd + cv * 4, is);
    load4f(coef + cv * 4, cA);
    load4f(coef + C + cv * 4, c2);
    load4f(coef + 2 * C + cv * 4, c3);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cB[k] = -cA[k] * c3[k] * is[k];
      cC[k] = -cA[k] * c2[k] - cB[k] * mu[k];
      cA[k] *= row_scale;
    }
  }
  for (;;) {
    const int w = dyn_next(sched);
    if (w >= q.V) break;
    long long u0, u1;
    dyn_range(q, w, u0, u1);
    const long long v1 = min(nvec, u1 * 1024);
#pragma unroll kFlatUnroll
    for (long long v = u0 * 1024 + threadIdx.x; v < v1; v += kDynThreads) {
      const uint2 raw = ldg8(y + v * 4);
      const long long gv = row_shift >= 0 ? ((v >> row_shift) * CV + cv) : v;
      float g[4], x[4], o[4];
      load_grad4<kGradF32>(dA_, gv, g);
      unpack4(raw, x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gk = fmaf(x[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
        o[k] = fmaf(cA[k], gk, fmaf(cB[k], x[k], cC[k]));
      }
      store4(dy + v * 4, o);
    }
  }
  dyn_retire(sched);
}

// ---- 2x2-pooled layers: a thread-iteration owns one pooled gradient vector (4 channels) and its 2x2 window of y.
// A worker owns a range of pooled rows (b, ho); its (row, vector) pairs are walked as one flat index so that two
// independent iterations (2 x (1 + 4) loads) are in flight whatever the row length.  ov_shift = log2(Wo * C/4).
template <bool kGradF32>
__global__ void __launch_bounds__(kDynThreads, 4)
bn_bwd_reduce_win4_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                          const float* __restrict__ scale, const float* __restrict__ shift, int H, int W, int C,
                          int ov_shift, DynPlan q, int G, float* __restrict__ partial, int* sched) {
  __shared__ float s_red[kDynThreads * 4];
  const int CV = C >> 2, Ho = H >> 1, cv = threadIdx.x & (CV - 1), pl = threadIdx.x / CV, lanes = kDynThreads / CV;
  const int cv_shift = __ffs(CV) - 1;
  const long long in_row = (long long)W * C;
  const int ov_mask = (1 << ov_shift) - 1;
  float sc[4], sh[4];
  load4f(scale + cv * 4, sc);
  load4f(shift + cv * 4, sh);
  for (;;) {
    const int w = dyn_next(sched);
    if (w >= q.V) break;
    long long u0, u1;
    dyn_range(q, w, u0, u1);
    const long long j1 = (u1 - u0) << ov_shift;
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sgy[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll kWinUnroll
    for (long long j = threadIdx.x; j < j1; j += kDynThreads) {
      const int row = (int)(u0 + (j >> ov_shift)), v = (int)j & ov_mask;
      const int b = row / Ho, ho = row - b * Ho, wo = v >> cv_shift;
      const __nv_bfloat16* src = y + ((long long)b * H + 2LL * ho) * in_row + ((long long)(2 * wo) * CV + cv) * 4;
      uint2 raw[4];
      raw[0] = ldg8(src);
      raw[1] = ldg8(src + C);
      raw[2] = ldg8(src + in_row);
      raw[3] = ldg8(src + in_row + C);
      float g[4];
      load_grad4<kGradF32>(dA_, ((long long)row << ov_shift) + v, g);
#pragma unroll
      for (int k = 0; k < 4; ++k) g[k] *= 0.25f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x[4];
        unpack4(raw[i], x);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float gk = fmaf(x[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
          sg[k] += gk;
          sgy[k] = fmaf(gk, x[k], sgy[k]);
        }
      }
    }
    dyn_write_partial(s_red, sg, sgy, C, cv, pl, lanes, partial + ((long long)G + w) * 2 * C);
    dyn_group_reduce(partial, w, q.V, G, C, sched + 2);
  }
  dyn_retire(sched);
}

template <bool kGradF32>
__global__ void __launch_bounds__(kDynThreads, 4)
bn_bwd_apply_win4_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ coef, int B, int H, int W, int C, int ov_shift, DynPlan q,
                         __nv_bfloat16* __restrict__ dy, int* sched) {
  const int CV = C >> 2, Ho = H >> 1, cv = threadIdx.x & (CV - 1);
  const int cv_shift = __ffs(CV) - 1;
  const long long in_row = (long long)W * C;
  const int ov_mask = (1 << ov_shift) - 1;
  float sc[4], sh[4], cA[4], cB[4], cC[4];
  {
    float mu[4], is[4], c1[4], c2[4], c3[4];
    load4f(scale + cv * 4, sc);
    load4f(shift + cv * 4, sh);
    load4f(mean + cv * 4, mu);
    load4f(invstd + cv * 4, is);
    load4f(coef + cv * 4, c1);
    load4f(coef + C + cv * 4, c2);
    load4f(coef + 2 * C + cv * 4, c3);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cA[k] = c1[k] * 0.25f;
      cB[k] = -c1[k] * c3[k] * is[k];
      cC[k] = -c1[k] * c2[k] - cB[k] * mu[k];
    }
  }
  for (;;) {
    const int w = dyn_next(sched);
    if (w >= q.V) break;
    long long u0, u1;
    dyn_range(q, w, u0, u1);
    const long long j1 = (u1 - u0) << ov_shift;
#pragma unroll kWinUnroll
    for (long long j = threadIdx.x; j < j1; j += kDynThreads) {
      const int row = (int)(u0 + (j >> ov_shift)), v = (int)j & ov_mask;
      const int b = row / Ho, ho = row - b * Ho, wo = v >> cv_shift;
      const long long x0 = ((long long)b * H + 2LL * ho) * in_row + ((long long)(2 * wo) * CV + cv) * 4;
      uint2 raw[4];
      raw[0] = ldg8(y + x0);
      raw[1] = ldg8(y + x0 + C);
      raw[2] = ldg8(y + x0 + in_row);
      raw[3] = ldg8(y + x0 + in_row + C);
      float g[4];
      load_grad4<kGradF32>(dA_, ((long long)row << ov_shift) + v, g);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x[4], o[4];
        unpack4(raw[i], x);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float gk = fmaf(x[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
          o[k] = fmaf(cA[k], gk, fmaf(cB[k], x[k], cC[k]));
        }
        store4(dy + x0 + (i >> 1) * in_row + (i & 1) * C, o);
      }
    }
    if (H & 1) {                                     // floor-mode tail row: no pooled gradient reaches it
      const int in_vecs = W * CV;
      for (int b = w; b < B; b += q.V) {
        const long long base = ((long long)b * H + (H - 1)) * in_row;
        for (int v = threadIdx.x; v < in_vecs; v += kDynThreads) {
          float x[4], o[4];
          unpack4(ldg8(y + base + (long long)v * 4), x);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = fmaf(cB[k], x[k], cC[k]);
          store4(dy + base + (long long)v * 4, o);
        }
      }
    }
  }
  dyn_retire(sched);
}

bool rows_path_ok(int H, int W, int C, int ph, int pw) {
  const int CV = C / 8;
  if (C % 8 != 0 || CV < 1 || CV > kRowThreads || kRowThreads % CV != 0) return false;
  if (ph == 1 && pw == 1) return true;
  if (ph == 1 && pw == W) return ((W * CV) & (W * CV - 1)) == 0;      // whole-row average (block 4 + mean over mel)
  return ph == 2 && pw == 2 && W % 2 == 0 && H >= 2;
}
int ilog2(long long n) {
  int s = 0;
  while ((1LL << s) < n) ++s;
  return s;
}
bool pow2(long long n) { return n > 0 && (n & (n - 1)) == 0; }

// ---- plan of the dynamically scheduled backward kernels ----------------------------------------------------------
struct BwdPlan {
  int kind;          // 0 generic kernels, 1 flat (un-pooled / whole-row pooled), 2 2x2 window
  int row_shift;     // flat: log2(W * C/4) for whole-row pooling, -1 un-pooled;  win: log2(Wo * C/4)
  float row_scale;
  long long nvec;    // flat: four-channel vectors
  DynPlan q;
  int grid;
  int G;             // rows the finalize kernel reads (fast paths: worker groups; generic: one per CTA)
};
BwdPlan make_bwd_plan(int B, int H, int W, int C, int ph, int pw) {
  BwdPlan p;
  memset(&p, 0, sizeof(p));
  const int CV = C / 4;
  const bool c_ok = C % 4 == 0 && CV >= 1 && CV <= kDynThreads && pow2(CV) && C <= kDynThreads * 4;
  const long long rows = (long long)B * H;
  long long unit_bytes = 0;
  if (c_ok && rows < (1LL << 31) && ph == 1 && (pw == 1 || (pw == W && pow2((long long)W * CV)))) {
    p.kind = 1;
    p.row_shift = pw == 1 ? -1 : ilog2((long long)W * CV);
    p.row_scale = pw == 1 ? 1.0f : 1.0f / (float)pw;
    p.nvec = rows * W * CV;
    p.q.n_units = (p.nvec + 1023) / 1024;
    unit_bytes = 1024 * 8;
  } else if (c_ok && rows < (1LL << 31) && ph == 2 && pw == 2 && W % 2 == 0 && H >= 2 && pow2((long long)(W / 2) * CV)) {
    p.kind = 2;
    p.row_shift = ilog2((long long)(W / 2) * CV);
    p.q.n_units = (long long)B * (H / 2);
    unit_bytes = 2LL * W * C * 2;
  } else {
    p.q.V = sm_count() * 3;          // rows of the partial workspace = grid of the generic reduce kernel
    p.grid = p.q.V;
    p.G = p.q.V;
    return p;
  }
  // workers of ~768 KB of y (never fewer than 4 per SM, never more than 16 per SM in the first tier); the last quarter
  // of the tensor goes to workers a quarter of that size
  const long long n = p.q.n_units;
  long long big = (768 * 1024) / unit_bytes;
  const long long lo = (n + sm_count() * 16LL - 1) / (sm_count() * 16LL), hi = n / (sm_count() * 4LL);
  if (big > hi) big = hi;
  if (big < lo) big = lo;
  if (big < 1) big = 1;
  long long small = big / 4;
  if (small < 1) small = 1;
  long long v_big = (n - n / 4) / big;
  const long long rest = n - v_big * big;
  const long long v_small = (rest + small - 1) / small;
  p.q.big = big; p.q.small = small; p.q.v_big = (int)v_big; p.q.V = (int)(v_big + v_small);
  // THREE CTAs per SM although four would fit on an empty SM (64 registers): four fill the register file, and a
  // weight-gradient CTA that becomes ready while they run could not be placed until the whole (persistent) grid retires.
  // With three, 16 K registers and > 210 KB of shared memory stay free on every SM at all times.
  p.grid = sm_count() * 3;
  if (p.grid > p.q.V) p.grid = p.q.V;
  p.G = (p.q.V + kGroup - 1) / kGroup;
  return p;
}
// log2(W * C/8) for the whole-row pooling form of the flat kernels, -1 for the un-pooled form
int row_pool_shift(int W, int C, int pw) {
  if (pw == 1) return -1;
  int n = W * (C / 8), s = 0;
  while ((1 << s) < n) ++s;
  return s;
}
// one full wave: `per_sm` CTAs of 256 threads are co-resident per SM (set by __launch_bounds__), so every
// CTA walks the same number of rows and there is no partial second wave
int rows_grid(int rows, int per_sm) {
  const int cap = sm_count() * per_sm;
  return rows < cap ? (rows < 1 ? 1 : rows) : cap;
}

int ew_grid(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_bn_finalize(const float* partial, int P, int C, double count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* scale, float* shift, float* save_mean,
                    float* save_invstd, sed_stream_t stream) {
  SED_REQUIRE(partial && scale && shift && P >= 1 && C >= 1 && count >= 1.0, "sed_bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(C, 32), dim3(32, kColLanes), 0, (cudaStream_t)stream>>>(
      partial, P, C, count, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, scale, shift,
      save_mean, save_invstd);
  SED_LAUNCH_CHECK("bn_finalize_kernel");
  return 0;
}

int sed_bn_eval_affine(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                       float eps, int C, float* scale, float* shift, float* save_mean, float* save_invstd,
                       sed_stream_t stream) {
  SED_REQUIRE(running_mean && running_var && scale && shift && C >= 1, "sed_bn_eval_affine: bad arguments");
  bn_eval_affine_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, gamma, beta,
                                                                            eps, C, scale, shift, save_mean, save_invstd);
  SED_LAUNCH_CHECK("bn_eval_affine_kernel");
  return 0;
}

int sed_bn_relu_pool_fwd(const void* y, const float* scale, const float* shift, int B, int H, int W, int C, int ph,
                         int pw, void* out, int out_is_f32, sed_stream_t stream) {
  SED_REQUIRE(y && scale && shift && out, "sed_bn_relu_pool_fwd: null pointer");
  SED_REQUIRE(C % 8 == 0 && ph >= 1 && pw >= 1 && H / ph >= 1 && W / pw >= 1, "sed_bn_relu_pool_fwd: bad shape");
  SED_REQUIRE(aligned(y, 16) && aligned(out, 16) && aligned(scale, 16) && aligned(shift, 16),
              "sed_bn_relu_pool_fwd: pointers must be 16-byte aligned");
  if (B == 0) return 0;
  const long long total = (long long)B * (H / ph) * (W / pw) * (C / 8);
  const int grid = ew_grid(total, 256);
  const __nv_bfloat16* yy = reinterpret_cast<const __nv_bfloat16*>(y);
  if (!out_is_f32 && ph == pw && rows_path_ok(H, W, C, ph, pw) && (long long)B * H < (1LL << 31)) {
    __nv_bfloat16* oo = reinterpret_cast<__nv_bfloat16*>(out);
    const int g = rows_grid(B * (H / ph), 4);
    if (ph == 1) bn_relu_fwd_flat_kernel<<<sm_count() * 4, kRowThreads, 0, (cudaStream_t)stream>>>(yy, scale, shift, (long long)B * H * W * (C / 8), C, oo);
    else bn_relu_pool_fwd_rows_kernel<2><<<g, kRowThreads, 0, (cudaStream_t)stream>>>(yy, scale, shift, B, H, W, C, oo);
    SED_LAUNCH_CHECK("bn_relu_pool_fwd_kernel (flat / rows)");
    return 0;
  }
  if (out_is_f32)
    bn_relu_pool_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(yy, scale, shift, B, H, W, C, ph, pw, out);
  else
    bn_relu_pool_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(yy, scale, shift, B, H, W, C, ph, pw, out);
  SED_LAUNCH_CHECK("bn_relu_pool_fwd_kernel");
  return 0;
}

int sed_bn_bwd_partials(int B, int H, int W, int C, int ph, int pw) {   // rows sed_bn_bwd_finalize reads (the FIRST rows)
  if (B < 1 || H < 1 || W < 1 || C < 8) return 0;
  return make_bwd_plan(B, H, W, C, ph, pw).G;
}
int sed_bn_bwd_workspace_rows(int B, int H, int W, int C, int ph, int pw) {   // rows of [2][C] floats the reduce pass needs
  if (B < 1 || H < 1 || W < 1 || C < 8) return 0;
  const BwdPlan p = make_bwd_plan(B, H, W, C, ph, pw);
  return p.kind == 0 ? p.G : p.G + p.q.V;
}
int sed_bn_bwd_sched_words(void) { return 2 + kMaxGroups; }

int sed_bn_relu_pool_bwd_reduce(const void* y, const void* dA, int grad_is_f32, const float* scale,
                                const float* shift, int B, int H, int W, int C, int ph, int pw, float* partial,
                                int* sched, sed_stream_t stream) {
  SED_REQUIRE(y && dA && scale && shift && partial && sched, "sed_bn_relu_pool_bwd_reduce: null pointer");
  SED_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "sed_bn_relu_pool_bwd_reduce: C=%d unsupported", C);
  SED_REQUIRE(B >= 1 && H / ph >= 1 && W / pw >= 1, "sed_bn_relu_pool_bwd_reduce: bad shape");
  SED_REQUIRE(aligned(y, 16) && aligned(dA, 16) && aligned(scale, 16) && aligned(shift, 16),
              "sed_bn_relu_pool_bwd_reduce: pointers must be 16-byte aligned");
  const BwdPlan p = make_bwd_plan(B, H, W, C, ph, pw);
  SED_REQUIRE(p.kind == 0 || p.G <= kMaxGroups, "sed_bn_relu_pool_bwd_reduce: %d worker groups exceed the sched workspace", p.G);
  const __nv_bfloat16* yy = reinterpret_cast<const __nv_bfloat16*>(y);
  cudaStream_t st = (cudaStream_t)stream;
  if (p.kind == 1) {
    if (grad_is_f32) bn_bwd_reduce_flat4_kernel<true><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, p.nvec, C, p.row_shift, p.row_scale, p.q, p.G, partial, sched);
    else bn_bwd_reduce_flat4_kernel<false><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, p.nvec, C, p.row_shift, p.row_scale, p.q, p.G, partial, sched);
    SED_LAUNCH_CHECK("bn_bwd_reduce_flat4_kernel");
    return 0;
  }
  if (p.kind == 2) {
    if (grad_is_f32) bn_bwd_reduce_win4_kernel<true><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, H, W, C, p.row_shift, p.q, p.G, partial, sched);
    else bn_bwd_reduce_win4_kernel<false><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, H, W, C, p.row_shift, p.q, p.G, partial, sched);
    SED_LAUNCH_CHECK("bn_bwd_reduce_win4_kernel");
    return 0;
  }
  const int lanes = 256 / (C / 8);
  const size_t smem = (size_t)lanes * 2 * C * sizeof(float);
  if (grad_is_f32)
    bn_relu_pool_bwd_reduce_kernel<true><<<p.grid, 256, smem, st>>>(yy, dA, scale, shift, B, H, W, C, ph, pw, partial);
  else
    bn_relu_pool_bwd_reduce_kernel<false><<<p.grid, 256, smem, st>>>(yy, dA, scale, shift, B, H, W, C, ph, pw, partial);
  SED_LAUNCH_CHECK("bn_relu_pool_bwd_reduce_kernel");
  return 0;
}

int sed_bn_bwd_finalize(const float* partial, int P, int C, double count, const float* gamma, const float* invstd,
                        const float* mean_for_gy, float* dgamma, float* dbeta, int accumulate, int frozen_stats,
                        float* coef, sed_stream_t stream) {
  SED_REQUIRE(partial && invstd && P >= 1 && C >= 1, "sed_bn_bwd_finalize: bad arguments");
  bn_bwd_finalize_kernel<<<ceil_div(C, 32), dim3(32, kColLanes), 0, (cudaStream_t)stream>>>(partial, P, C, count, gamma, invstd,
                                                                             mean_for_gy, dgamma, dbeta, accumulate,
                                                                             frozen_stats, coef);
  SED_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  return 0;
}

int sed_bn_relu_pool_bwd_apply(const void* y, const void* dA, int grad_is_f32, const float* scale, const float* shift,
                               const float* mean, const float* invstd, const float* coef, int B, int H, int W, int C,
                               int ph, int pw, void* dy, int* sched, sed_stream_t stream) {
  SED_REQUIRE(y && dA && scale && shift && mean && invstd && coef && dy && sched,
              "sed_bn_relu_pool_bwd_apply: null pointer");
  SED_REQUIRE(C % 8 == 0, "sed_bn_relu_pool_bwd_apply: C must be a multiple of 8");
  SED_REQUIRE(aligned(y, 16) && aligned(dA, 16) && aligned(dy, 16) && aligned(coef, 16),
              "sed_bn_relu_pool_bwd_apply: pointers must be 16-byte aligned");
  if (B == 0) return 0;
  const BwdPlan p = make_bwd_plan(B, H, W, C, ph, pw);
  const __nv_bfloat16* yy = reinterpret_cast<const __nv_bfloat16*>(y);
  __nv_bfloat16* dd = reinterpret_cast<__nv_bfloat16*>(dy);
  cudaStream_t st = (cudaStream_t)stream;
  if (p.kind == 1) {
    if (grad_is_f32) bn_bwd_apply_flat4_kernel<true><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, p.nvec, C, p.row_shift, p.row_scale, p.q, dd, sched);
    else bn_bwd_apply_flat4_kernel<false><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, p.nvec, C, p.row_shift, p.row_scale, p.q, dd, sched);
    SED_LAUNCH_CHECK("bn_bwd_apply_flat4_kernel");
    return 0;
  }
  if (p.kind == 2) {
    if (grad_is_f32) bn_bwd_apply_win4_kernel<true><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, B, H, W, C, p.row_shift, p.q, dd, sched);
    else bn_bwd_apply_win4_kernel<false><<<p.grid, kDynThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, B, H, W, C, p.row_shift, p.q, dd, sched);
    SED_LAUNCH_CHECK("bn_bwd_apply_win4_kernel");
    return 0;
  }
  const long long total = (long long)B * H * W * (C / 8);
  const int grid = ew_grid(total, 256);
  if (grad_is_f32)
    bn_relu_pool_bwd_apply_kernel<true><<<grid, 256, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, B, H, W, C, ph, pw, dd);
  else
    bn_relu_pool_bwd_apply_kernel<false><<<grid, 256, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, B, H, W, C, ph, pw, dd);
  SED_LAUNCH_CHECK("bn_relu_pool_bwd_apply_kernel");
  return 0;
}

}  // extern "C"
