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
                                      int C, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.0f / sqrtf(running_var[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - running_mean[c] * g * invstd;
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
                                               const float* __restrict__ mean, const float* __restrict__ invstd,
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
  float sc[8], sh[8], mu[8], is[8], sg[8], sgx[8];
  load8f(scale + cv * kVec, sc);
  load8f(shift + cv * kVec, sh);
  load8f(mean + cv * kVec, mu);
  load8f(invstd + cv * kVec, is);
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
        sgx[k] += gk * (v[k] - mu[k]) * is[k];
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
    coef[C + c] = (float)(s / count);
    coef[2 * C + c] = (float)(sx / count);
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

template <bool kGradF32>
__global__ void __launch_bounds__(kRowThreads, 3)
bn_bwd_reduce_flat_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd, long long nvec, int C,
                          int row_shift, float row_scale, float* __restrict__ partial) {
  // row_shift >= 0: "pool over the whole row" (block 4: 1 x W average = torch.mean(x, dim=3)): dA is (rows, C) and
  // vector v of y belongs to row v >> row_shift (W * C/8 is a power of two), g = dA / W
  extern __shared__ float s_red[];                 // [lanes][2*C]
  const int CV = C >> 3, cv = threadIdx.x % CV, pl = threadIdx.x / CV, lanes = kRowThreads / CV;
  float sc[8], sh[8], mu[8], is[8], sg[8], sgx[8];
  load8f(scale + cv * 8, sc);
  load8f(shift + cv * 8, sh);
  load8f(mean + cv * 8, mu);
  load8f(invstd + cv * 8, is);
#pragma unroll
  for (int k = 0; k < 8; ++k) sg[k] = sgx[k] = 0.f;
  const long long stride = (long long)gridDim.x * kRowThreads;
#pragma unroll 4
  for (long long v = (long long)blockIdx.x * kRowThreads + threadIdx.x; v < nvec; v += stride) {
    float g[8], x[8];
    const uint4 raw = ldg16(y + v * 8);
    const long long gv = row_shift >= 0 ? ((v >> row_shift) * CV + cv) : v;
    if (kGradF32) load8f(reinterpret_cast<const float*>(dA_) + gv * 8, g);
    else unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(dA_) + gv * 8), g);
    unpack8(raw, x);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gk = fmaf(x[k], sc[k], sh[k]) > 0.f ? g[k] * row_scale : 0.f;
      sg[k] += gk;
      sgx[k] += gk * (x[k] - mu[k]) * is[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s_red[pl * 2 * C + cv * 8 + k] = sg[k];
    s_red[pl * 2 * C + C + cv * 8 + k] = sgx[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kRowThreads) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * 2 * C + i];
    partial[(long long)blockIdx.x * 2 * C + i] = a;
  }
}

template <bool kGradF32>
__global__ void __launch_bounds__(kRowThreads, 3)
bn_bwd_apply_flat_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ coef, long long nvec, int C, int row_shift, float row_scale,
                         __nv_bfloat16* __restrict__ dy) {
  const int CV = C >> 3, cv = threadIdx.x % CV;
  float sc[8], sh[8], cA[8], cB[8], cC[8];
  {
    float mu[8], is[8], c2[8], c3[8];
    load8f(scale + cv * 8, sc);
    load8f(shift + cv * 8, sh);
    load8f(mean + cv * 8, mu);
    load8f(invstd + cv * 8, is);
    load8f(coef + cv * 8, cA);
    load8f(coef + C + cv * 8, c2);
    load8f(coef + 2 * C + cv * 8, c3);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cB[k] = -cA[k] * c3[k] * is[k];
      cC[k] = -cA[k] * c2[k] - cB[k] * mu[k];
      cA[k] *= row_scale;
    }
  }
  const long long stride = (long long)gridDim.x * kRowThreads;
#pragma unroll 4
  for (long long v = (long long)blockIdx.x * kRowThreads + threadIdx.x; v < nvec; v += stride) {
    float g[8], x[8], o[8];
    const uint4 raw = ldg16(y + v * 8);
    const long long gv = row_shift >= 0 ? ((v >> row_shift) * CV + cv) : v;
    if (kGradF32) load8f(reinterpret_cast<const float*>(dA_) + gv * 8, g);
    else unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(dA_) + gv * 8), g);
    unpack8(raw, x);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gk = fmaf(x[k], sc[k], sh[k]) > 0.f ? g[k] : 0.f;
      o[k] = fmaf(cA[k], gk, fmaf(cB[k], x[k], cC[k]));
    }
    store8(dy + v * 8, o);
  }
}

// ---- 2x2-pool backward, window-structured: a thread-iteration owns one pooled output vector (8 channels) and
// its 2x2 window of y: ONE dA load + FOUR y loads in flight (instead of 4 x (1 + 1) with the dA vector fetched by
// four different threads), no integer division, and the floor-mode tail row (odd H) is handled separately.
template <bool kGradF32>
__global__ void __launch_bounds__(kRowThreads, 3)
bn_bwd_reduce_win2_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ mean, const float* __restrict__ invstd, int B, int H, int W,
                          int C, float* __restrict__ partial) {
  extern __shared__ float s_red[];                 // [lanes][2*C]
  const int CV = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV, lanes = kRowThreads / CV, cv_shift = __ffs(CV) - 1;
  float sc[8], sh[8], mu[8], is[8], sg[8], sgx[8];
  load8f(scale + cv * 8, sc);
  load8f(shift + cv * 8, sh);
  load8f(mean + cv * 8, mu);
  load8f(invstd + cv * 8, is);
#pragma unroll
  for (int k = 0; k < 8; ++k) sg[k] = sgx[k] = 0.f;
  const int out_vecs = Wo * CV;
  const long long in_row = (long long)W * C;
  const int rows = B * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / Ho, ho = row - b * Ho;
    const __nv_bfloat16* src = y + ((long long)b * H + 2LL * ho) * in_row;
    const long long drow = (long long)row * Wo * C;
#pragma unroll 1
    for (int v = threadIdx.x; v < out_vecs; v += kRowThreads) {
      const int wo = v >> cv_shift;
      const long long x0 = ((long long)(2 * wo) * CV + cv) * 8;
      uint4 raw[4];
      raw[0] = ldg16(src + x0);
      raw[1] = ldg16(src + x0 + C);
      raw[2] = ldg16(src + in_row + x0);
      raw[3] = ldg16(src + in_row + x0 + C);
      float g[8];
      if (kGradF32) load8f(reinterpret_cast<const float*>(dA_) + drow + (long long)v * 8, g);
      else { const uint4 graw = ldg16(reinterpret_cast<const __nv_bfloat16*>(dA_) + drow + (long long)v * 8); unpack8(graw, g); }
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] *= 0.25f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x[8];
        unpack8(raw[i], x);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float act = fmaf(x[k], sc[k], sh[k]);
          const float gk = act > 0.f ? g[k] : 0.f;
          sg[k] += gk;
          sgx[k] += gk * (x[k] - mu[k]) * is[k];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s_red[pl * 2 * C + cv * 8 + k] = sg[k];
    s_red[pl * 2 * C + C + cv * 8 + k] = sgx[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kRowThreads) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * 2 * C + i];
    partial[(long long)blockIdx.x * 2 * C + i] = a;
  }
}

template <bool kGradF32>
__global__ void __launch_bounds__(kRowThreads, 3)
bn_bwd_apply_win2_kernel(const __nv_bfloat16* __restrict__ y, const void* __restrict__ dA_,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ coef, int B, int H, int W, int C,
                         __nv_bfloat16* __restrict__ dy) {
  const int CV = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int cv = threadIdx.x % CV, cv_shift = __ffs(CV) - 1;
  float sc[8], sh[8], cA[8], cB[8], cC[8];
  {
    float mu[8], is[8], c1[8], c2[8], c3[8];
    load8f(scale + cv * 8, sc);
    load8f(shift + cv * 8, sh);
    load8f(mean + cv * 8, mu);
    load8f(invstd + cv * 8, is);
    load8f(coef + cv * 8, c1);
    load8f(coef + C + cv * 8, c2);
    load8f(coef + 2 * C + cv * 8, c3);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cA[k] = c1[k] * 0.25f;
      cB[k] = -c1[k] * c3[k] * is[k];
      cC[k] = -c1[k] * c2[k] - cB[k] * mu[k];
    }
  }
  const int out_vecs = Wo * CV;
  const long long in_row = (long long)W * C;
  const int rows = B * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / Ho, ho = row - b * Ho;
    const long long base = ((long long)b * H + 2LL * ho) * in_row;
    const long long drow = (long long)row * Wo * C;
#pragma unroll 1
    for (int v = threadIdx.x; v < out_vecs; v += kRowThreads) {
      const int wo = v >> cv_shift;
      const long long x0 = base + ((long long)(2 * wo) * CV + cv) * 8;
      const long long offs[4] = {x0, x0 + C, x0 + in_row, x0 + in_row + C};
      uint4 raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) raw[i] = ldg16(y + offs[i]);
      float g[8];
      if (kGradF32) load8f(reinterpret_cast<const float*>(dA_) + drow + (long long)v * 8, g);
      else { const uint4 graw = ldg16(reinterpret_cast<const __nv_bfloat16*>(dA_) + drow + (long long)v * 8); unpack8(graw, g); }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x[8], o[8];
        unpack8(raw[i], x);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float act = fmaf(x[k], sc[k], sh[k]);
          const float gk = act > 0.f ? g[k] : 0.f;
          o[k] = fmaf(cA[k], gk, fmaf(cB[k], x[k], cC[k]));
        }
        store8(dy + offs[i], o);
      }
    }
  }
  if (H & 1) {                                       // floor-mode tail row: no pooled gradient reaches it
    const int in_vecs = W * CV;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      const long long base = ((long long)b * H + (H - 1)) * in_row;
      for (int v = threadIdx.x; v < in_vecs; v += kRowThreads) {
        float x[8], o[8];
        unpack8(ldg16(y + base + (long long)v * 8), x);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(cB[k], x[k], cC[k]);
        store8(dy + base + (long long)v * 8, o);
      }
    }
  }
}

bool rows_path_ok(int H, int W, int C, int ph, int pw) {
  const int CV = C / 8;
  if (C % 8 != 0 || CV < 1 || CV > kRowThreads || kRowThreads % CV != 0) return false;
  if (ph == 1 && pw == 1) return true;
  if (ph == 1 && pw == W) return ((W * CV) & (W * CV - 1)) == 0;      // whole-row average (block 4 + mean over mel)
  return ph == 2 && pw == 2 && W % 2 == 0 && H >= 2;
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
                       float eps, int C, float* scale, float* shift, sed_stream_t stream) {
  SED_REQUIRE(running_mean && running_var && scale && shift && C >= 1, "sed_bn_eval_affine: bad arguments");
  bn_eval_affine_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, gamma, beta,
                                                                            eps, C, scale, shift);
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

int sed_bn_bwd_partials(int C) {   // rows of the partial workspace used by the reduce pass
  (void)C;
  return sm_count() * 3;           // = one resident wave of the row-structured reduce kernel (3 CTAs / SM)
}

int sed_bn_relu_pool_bwd_reduce(const void* y, const void* dA, int grad_is_f32, const float* scale,
                                const float* shift, const float* mean, const float* invstd, int B, int H, int W,
                                int C, int ph, int pw, float* partial, sed_stream_t stream) {
  SED_REQUIRE(y && dA && scale && shift && mean && invstd && partial, "sed_bn_relu_pool_bwd_reduce: null pointer");
  SED_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "sed_bn_relu_pool_bwd_reduce: C=%d unsupported", C);
  const int grid = sed_bn_bwd_partials(C);
  const int lanes = 256 / (C / 8);
  const size_t smem = (size_t)lanes * 2 * C * sizeof(float);
  const __nv_bfloat16* yy = reinterpret_cast<const __nv_bfloat16*>(y);
  if (rows_path_ok(H, W, C, ph, pw) && (long long)B * H < (1LL << 31)) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = (long long)B * H * W * (C / 8);
    if (ph == 1) {
      const int rs = row_pool_shift(W, C, pw);
      const float rsc = pw == 1 ? 1.0f : 1.0f / (float)pw;
      if (grad_is_f32) bn_bwd_reduce_flat_kernel<true><<<grid, kRowThreads, smem, st>>>(yy, dA, scale, shift, mean, invstd, nvec, C, rs, rsc, partial);
      else bn_bwd_reduce_flat_kernel<false><<<grid, kRowThreads, smem, st>>>(yy, dA, scale, shift, mean, invstd, nvec, C, rs, rsc, partial);
    }
    else if (grad_is_f32) bn_bwd_reduce_win2_kernel<true><<<grid, kRowThreads, smem, st>>>(yy, dA, scale, shift, mean, invstd, B, H, W, C, partial);
    else bn_bwd_reduce_win2_kernel<false><<<grid, kRowThreads, smem, st>>>(yy, dA, scale, shift, mean, invstd, B, H, W, C, partial);
    SED_LAUNCH_CHECK("bn_bwd_reduce_kernel");
    return 0;
  }
  if (grad_is_f32)
    bn_relu_pool_bwd_reduce_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(yy, dA, scale, shift, mean, invstd,
                                                                                    B, H, W, C, ph, pw, partial);
  else
    bn_relu_pool_bwd_reduce_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(yy, dA, scale, shift, mean, invstd,
                                                                                     B, H, W, C, ph, pw, partial);
  SED_LAUNCH_CHECK("bn_relu_pool_bwd_reduce_kernel");
  return 0;
}

int sed_bn_bwd_finalize(const float* partial, int P, int C, double count, const float* gamma, const float* invstd,
                        const float* mean_for_gy, float* dgamma, float* dbeta, int accumulate, float* coef,
                        sed_stream_t stream) {
  SED_REQUIRE(partial && invstd && P >= 1 && C >= 1, "sed_bn_bwd_finalize: bad arguments");
  bn_bwd_finalize_kernel<<<ceil_div(C, 32), dim3(32, kColLanes), 0, (cudaStream_t)stream>>>(partial, P, C, count, gamma, invstd,
                                                                             mean_for_gy, dgamma, dbeta, accumulate, coef);
  SED_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  return 0;
}

int sed_bn_relu_pool_bwd_apply(const void* y, const void* dA, int grad_is_f32, const float* scale, const float* shift,
                               const float* mean, const float* invstd, const float* coef, int B, int H, int W, int C,
                               int ph, int pw, void* dy, sed_stream_t stream) {
  SED_REQUIRE(y && dA && scale && shift && mean && invstd && coef && dy, "sed_bn_relu_pool_bwd_apply: null pointer");
  SED_REQUIRE(C % 8 == 0, "sed_bn_relu_pool_bwd_apply: C must be a multiple of 8");
  if (B == 0) return 0;
  const long long total = (long long)B * H * W * (C / 8);
  const int grid = ew_grid(total, 256);
  const __nv_bfloat16* yy = reinterpret_cast<const __nv_bfloat16*>(y);
  __nv_bfloat16* dd = reinterpret_cast<__nv_bfloat16*>(dy);
  if (rows_path_ok(H, W, C, ph, pw) && (long long)B * H < (1LL << 31)) {
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = (long long)B * H * W * (C / 8);
    if (ph == 1) {
      const int gf = sm_count() * 3;
      const int rs = row_pool_shift(W, C, pw);
      const float rsc = pw == 1 ? 1.0f : 1.0f / (float)pw;
      if (grad_is_f32) bn_bwd_apply_flat_kernel<true><<<gf, kRowThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, nvec, C, rs, rsc, dd);
      else bn_bwd_apply_flat_kernel<false><<<gf, kRowThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, nvec, C, rs, rsc, dd);
    }
    else {
      const int g2 = rows_grid(B * (H / 2), 3);
      if (grad_is_f32) bn_bwd_apply_win2_kernel<true><<<g2, kRowThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, B, H, W, C, dd);
      else bn_bwd_apply_win2_kernel<false><<<g2, kRowThreads, 0, st>>>(yy, dA, scale, shift, mean, invstd, coef, B, H, W, C, dd);
    }
    SED_LAUNCH_CHECK("bn_bwd_apply_kernel");
    return 0;
  }
  if (grad_is_f32)
    bn_relu_pool_bwd_apply_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(yy, dA, scale, shift, mean, invstd,
                                                                                coef, B, H, W, C, ph, pw, dd);
  else
    bn_relu_pool_bwd_apply_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(yy, dA, scale, shift, mean, invstd,
                                                                                 coef, B, H, W, C, ph, pw, dd);
  SED_LAUNCH_CHECK("bn_relu_pool_bwd_apply_kernel");
  return 0;
}

}  // extern "C"
