// heads.cu -- classifier heads and the loss: small-N (N = classes_num = 17) work that does not
// fill an MMA tile, done with warp-shuffle / CTA-reduction kernels in fp32.
//
// Replaces, relative to /root/reference/pytorch:
//   nn.Linear(512, 17) + sigmoid (models.py:306-307 and siblings)      -> sed_linear_small_fwd/bwd
//   Conv1d(k=1) att / cla of AttBlock (models.py:123-124, :137, :141)  -> sed_linear_small_fwd/bwd
//   interpolate x8 + mean / max over time (models.py:58-69, :227, :312) -> sed_head_pool_fwd/bwd
//   AttBlock.forward clamp/exp/normalise/sigmoid/weighted sum (:135-143) -> sed_head_att_fwd/bwd
//   F.binary_cross_entropy (losses.py:12)                                -> sed_bce_fwd_bwd
// Features arrive time-major (B, T, C) fp32 (that is x.transpose(1, 2) of the reference).
#include "common.cuh"

namespace sed {
namespace {

constexpr int kMaxK = 32;

// out[r][k] = bias[k] + sum_c x[r][c] * W[k][c]      one warp per row
__global__ void linear_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                        const float* __restrict__ bias, long long R, int C, int K,
                                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < R; r += nwarps) {
    float acc[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) acc[k] = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xv = x[r * C + c];
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        if (k < K) acc[k] = fmaf(xv, __ldg(W + (long long)k * C + c), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < kMaxK; ++k)
      if (k < K) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) out[r * K + k] = v + (bias ? bias[k] : 0.f);
      }
  }
}

// dx[r][c] = sum_k dout[r][k] * W[k][c]
__global__ void linear_small_dx_kernel(const float* __restrict__ dout, const float* __restrict__ W, long long R, int C,
                                       int K, float* __restrict__ dx, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < R; r += nwarps) {
    float d[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) d[k] = (k < K) ? dout[r * K + k] : 0.f;
    for (int c = lane; c < C; c += 32) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        if (k < K) a = fmaf(d[k], __ldg(W + (long long)k * C + c), a);
      dx[r * C + c] = accumulate ? dx[r * C + c] + a : a;
    }
  }
}

// partial[blk][k][c] = sum over the block's rows of dout[r][k] * x[r][c];  partial_b[blk][k] = sum dout[r][k]
__global__ void linear_small_dw_kernel(const float* __restrict__ dout, const float* __restrict__ x, long long R, int C,
                                       int K, float* __restrict__ partial_w, float* __restrict__ partial_b) {
  extern __shared__ float s_d[];                       // [rows_chunk][K]
  const int chunk = 32;
  const long long per = (R + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(R, r0 + per);
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    float acc[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) acc[k] = 0.f;
    for (long long rb = r0; rb < r1; rb += chunk) {
      const int n = (int)min((long long)chunk, r1 - rb);
      __syncthreads();
      for (int i = threadIdx.x; i < n * K; i += blockDim.x) s_d[i] = dout[rb * K + i];
      __syncthreads();
      if (c < C)
        for (int j = 0; j < n; ++j) {
          const float xv = x[(rb + j) * C + c];
#pragma unroll
          for (int k = 0; k < kMaxK; ++k)
            if (k < K) acc[k] = fmaf(s_d[j * K + k], xv, acc[k]);
        }
    }
    if (c < C)
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        if (k < K) partial_w[((long long)blockIdx.x * K + k) * C + c] = acc[k];
  }
  if (partial_b && threadIdx.x < K) {
    float a = 0.f;
    for (long long r = r0; r < r1; ++r) a += dout[r * K + threadIdx.x];
    partial_b[(long long)blockIdx.x * K + threadIdx.x] = a;
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// FrameAvg / FrameMax: prob = sigmoid(logit); frame = repeat(prob, ratio); clip = mean_t / max_t.
// one block per clip.  mode 0 = mean, 1 = max.
__global__ void head_pool_fwd_kernel(const float* __restrict__ logit, int T, int K, int ratio, int mode,
                                     float* __restrict__ prob, float* __restrict__ clip, int* __restrict__ argmax,
                                     float* __restrict__ frame) {
  const int b = blockIdx.x;
  const float* lg = logit + (long long)b * T * K;
  float* pr = prob + (long long)b * T * K;
  for (int i = threadIdx.x; i < T * K; i += blockDim.x) pr[i] = sigmoidf_(lg[i]);
  __syncthreads();
  if (frame) {
    float* fr = frame + (long long)b * T * ratio * K;
    for (int i = threadIdx.x; i < T * ratio * K; i += blockDim.x) {
      const int k = i % K, tt = i / K;
      fr[i] = pr[(tt / ratio) * K + k];               // exact index map t_out -> t_out / ratio
    }
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (mode == 0) {
      float a = 0.f;
      for (int t = 0; t < T; ++t) a += pr[t * K + k];
      clip[b * K + k] = a / (float)T;
    } else {
      float m = pr[k];
      int am = 0;
      for (int t = 1; t < T; ++t)
        if (pr[t * K + k] > m) { m = pr[t * K + k]; am = t; }
      clip[b * K + k] = m;
      if (argmax) argmax[b * K + k] = am;
    }
  }
}

__global__ void head_pool_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ dclip,
                                     const int* __restrict__ argmax, int T, int K, int mode,
                                     float* __restrict__ dlogit) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < T * K; i += blockDim.x) {
    const int k = i % K, t = i / K;
    const float p = prob[(long long)b * T * K + i];
    float g;
    if (mode == 0) g = dclip[b * K + k] / (float)T;
    else g = (argmax[b * K + k] == t) ? dclip[b * K + k] : 0.f;
    dlogit[(long long)b * T * K + i] = g * p * (1.f - p);
  }
}

// AttBlock tail.  one block per clip; threads over classes for the time reductions.
__global__ void head_att_fwd_kernel(const float* __restrict__ att_logit, const float* __restrict__ cla_logit, int T,
                                    int K, int ratio, int sigmoid_act, float temperature,
                                    float* __restrict__ norm_att /* (B,K,T) */, float* __restrict__ cla /* (B,K,T) */,
                                    float* __restrict__ clip, float* __restrict__ frame /* (B,T*ratio,K) */) {
  const int b = blockIdx.x;
  const float* al = att_logit + (long long)b * T * K;
  const float* cl = cla_logit + (long long)b * T * K;
  float* na = norm_att + (long long)b * K * T;
  float* ca = cla + (long long)b * K * T;
  for (int i = threadIdx.x; i < T * K; i += blockDim.x) {
    const int k = i % K, t = i / K;
    const float a = fminf(fmaxf(al[i], -10.f), 10.f);
    na[k * T + t] = expf(a / temperature) + 1e-6f;
    ca[k * T + t] = sigmoid_act ? sigmoidf_(cl[i]) : cl[i];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += na[k * T + t];
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const float n = na[k * T + t] / s;
      na[k * T + t] = n;
      acc += n * ca[k * T + t];
    }
    clip[b * K + k] = acc;
  }
  __syncthreads();
  if (frame) {
    float* fr = frame + (long long)b * T * ratio * K;
    for (int i = threadIdx.x; i < T * ratio * K; i += blockDim.x) {
      const int k = i % K, tt = i / K;
      fr[i] = ca[k * T + tt / ratio];
    }
  }
}

__global__ void head_att_bwd_kernel(const float* __restrict__ att_logit, const float* __restrict__ norm_att,
                                    const float* __restrict__ cla, const float* __restrict__ clip,
                                    const float* __restrict__ dclip, int T, int K, int sigmoid_act, float temperature,
                                    float* __restrict__ d_att_logit, float* __restrict__ d_cla_logit) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < T * K; i += blockDim.x) {
    const int k = i % K, t = i / K;
    const float g = dclip[b * K + k];
    const float n = norm_att[((long long)b * K + k) * T + t];
    const float c = cla[((long long)b * K + k) * T + t];
    const float raw = att_logit[(long long)b * T * K + i];
    // clip = sum_t e_t c_t / S  =>  dclip/de_t = (c_t - clip)/S ;  e = exp(a/temp) + 1e-6 ;  n = e/S
    // d a_t = g * (c_t - clip) * (e_t - 1e-6) / (S * temp) = g * (c_t - clip) * (n_t - 1e-6/S) / temp
    // 1e-6/S is recovered from n and the clamped logit: e_t = exp(a_t/temp) + 1e-6, S = e_t / n_t.
    const float a = fminf(fmaxf(raw, -10.f), 10.f);
    const float ex = expf(a / temperature);
    const float S = (ex + 1e-6f) / n;
    float da = g * (c - clip[b * K + k]) * ex / (S * temperature);
    if (raw < -10.f || raw > 10.f) da = 0.f;
    d_att_logit[(long long)b * T * K + i] = da;
    const float dc = g * n;
    d_cla_logit[(long long)b * T * K + i] = sigmoid_act ? dc * c * (1.f - c) : dc;
  }
}

// mean BCE with torch semantics (log clamped at -100; backward denominator clamped at 1e-12).
__global__ void bce_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n, float grad_scale,
                           float* __restrict__ loss, float* __restrict__ dp) {
  __shared__ double s_part[32];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float pi = p[i], ti = t[i];
    const float l1 = fmaxf(logf(pi), -100.f), l0 = fmaxf(log1pf(-pi), -100.f);
    acc += -(double)(ti * l1 + (1.f - ti) * l0);
    if (dp) dp[i] = grad_scale * (pi - ti) / fmaxf((1.f - pi) * pi, 1e-12f) / (float)n;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += s_part[w];
    *loss = (float)(a / (double)n);
  }
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_linear_partials(void) { return sm_count(); }

int sed_linear_small_fwd(const float* x, const float* W, const float* bias, long long R, int C, int K, float* out,
                         sed_stream_t stream) {
  SED_REQUIRE(x && W && out, "sed_linear_small_fwd: null pointer");
  SED_REQUIRE(K >= 1 && K <= kMaxK, "sed_linear_small_fwd: K=%d must be in [1, %d]", K, kMaxK);
  if (R == 0) return 0;
  const int grid = (int)min((R + 7) / 8, (long long)sm_count() * 8);
  linear_small_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, W, bias, R, C, K, out);
  SED_LAUNCH_CHECK("linear_small_fwd_kernel");
  return 0;
}

int sed_linear_small_bwd(const float* dout, const float* x, const float* W, long long R, int C, int K, float* dx,
                         int dx_accumulate, float* partial_w, float* partial_b, sed_stream_t stream) {
  SED_REQUIRE(dout && x && W, "sed_linear_small_bwd: null pointer");
  SED_REQUIRE(K >= 1 && K <= kMaxK, "sed_linear_small_bwd: K=%d must be in [1, %d]", K, kMaxK);
  if (R == 0) return 0;
  if (dx) {
    const int grid = (int)min((R + 7) / 8, (long long)sm_count() * 8);
    linear_small_dx_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dout, W, R, C, K, dx, dx_accumulate);
    SED_LAUNCH_CHECK("linear_small_dx_kernel");
  }
  if (partial_w) {
    linear_small_dw_kernel<<<sed_linear_partials(), 256, 32 * K * sizeof(float), (cudaStream_t)stream>>>(
        dout, x, R, C, K, partial_w, partial_b);
    SED_LAUNCH_CHECK("linear_small_dw_kernel");
  }
  return 0;
}

int sed_head_pool_fwd(const float* logit, int B, int T, int K, int ratio, int mode, float* prob, float* clip,
                      int* argmax, float* frame, sed_stream_t stream) {
  SED_REQUIRE(logit && prob && clip && (mode == 0 || (mode == 1 && argmax)), "sed_head_pool_fwd: bad arguments");
  if (B == 0) return 0;
  head_pool_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(logit, T, K, ratio, mode, prob, clip, argmax, frame);
  SED_LAUNCH_CHECK("head_pool_fwd_kernel");
  return 0;
}

int sed_head_pool_bwd(const float* prob, const float* dclip, const int* argmax, int B, int T, int K, int mode,
                      float* dlogit, sed_stream_t stream) {
  SED_REQUIRE(prob && dclip && dlogit && (mode == 0 || (mode == 1 && argmax)), "sed_head_pool_bwd: bad arguments");
  if (B == 0) return 0;
  head_pool_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(prob, dclip, argmax, T, K, mode, dlogit);
  SED_LAUNCH_CHECK("head_pool_bwd_kernel");
  return 0;
}

int sed_head_att_fwd(const float* att_logit, const float* cla_logit, int B, int T, int K, int ratio, int sigmoid_act,
                     float temperature, float* norm_att, float* cla, float* clip, float* frame, sed_stream_t stream) {
  SED_REQUIRE(att_logit && cla_logit && norm_att && cla && clip, "sed_head_att_fwd: null pointer");
  if (B == 0) return 0;
  head_att_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(att_logit, cla_logit, T, K, ratio, sigmoid_act, temperature,
                                                           norm_att, cla, clip, frame);
  SED_LAUNCH_CHECK("head_att_fwd_kernel");
  return 0;
}

int sed_head_att_bwd(const float* att_logit, const float* norm_att, const float* cla, const float* clip,
                     const float* dclip, int B, int T, int K, int sigmoid_act, float temperature, float* d_att_logit,
                     float* d_cla_logit, sed_stream_t stream) {
  SED_REQUIRE(att_logit && norm_att && cla && clip && dclip && d_att_logit && d_cla_logit,
              "sed_head_att_bwd: null pointer");
  if (B == 0) return 0;
  head_att_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(att_logit, norm_att, cla, clip, dclip, T, K, sigmoid_act,
                                                           temperature, d_att_logit, d_cla_logit);
  SED_LAUNCH_CHECK("head_att_bwd_kernel");
  return 0;
}

int sed_bce_fwd_bwd(const float* prob, const float* target, long long n, float grad_scale, float* loss, float* dprob,
                    sed_stream_t stream) {
  SED_REQUIRE(prob && target && loss && n >= 1, "sed_bce_fwd_bwd: bad arguments");
  bce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(prob, target, n, grad_scale, loss, dprob);
  SED_LAUNCH_CHECK("bce_kernel");
  return 0;
}

}  // extern "C"
