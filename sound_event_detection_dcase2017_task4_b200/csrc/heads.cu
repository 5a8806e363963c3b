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
constexpr int kMaxPairK = 40;             // K1 + K2 of the paired forward (10 float4 accumulators per thread; 11 spill)

constexpr int kMaxKG = kMaxK / 4;         // K padded to float4 groups
constexpr int kFwdRows = 64;              // rows per block of the forward kernel
constexpr int kFwdChunk = 32;             // columns of x staged per step

// out[r][k] = bias[k] + sum_c x[r][c] * W[k][c]
// block = 128 threads = 64 rows x 2 halves of the K groups.  W^T is staged once per block as [c][Kp]
// (Kp = K padded to a multiple of 4) so that a thread reads its weights as warp-broadcast float4s; x is
// staged 64 rows x 32 columns at a time (coalesced 128 B row segments, padded rows: conflict-free).
__global__ void __launch_bounds__(128)
linear_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                        long long R, int C, int K, float* __restrict__ out) {
  extern __shared__ float4 s_lin4[];
  const int KG = (K + 3) >> 2, Kp = KG * 4;
  float* s_w = reinterpret_cast<float*>(s_lin4);                 // [C][Kp]
  float* s_x = s_w + (size_t)C * Kp;                             // [64][33]
  for (int i = threadIdx.x; i < C * Kp; i += blockDim.x) {       // coalesced rows of W, transposed into shared memory
    const int k = i / C, c = i - k * C;                          // (reading down a column of W costs one sector per value)
    s_w[c * Kp + k] = k < K ? W[(long long)k * C + c] : 0.f;
  }
  const int r_loc = threadIdx.x & 63, kh = threadIdx.x >> 6;     // kh is warp-uniform
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r0 = (long long)blockIdx.x * kFwdRows;
  float4 acc[(kMaxKG + 1) / 2];
#pragma unroll
  for (int j = 0; j < (kMaxKG + 1) / 2; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* s_w4 = reinterpret_cast<const float4*>(s_w);
  for (int c0 = 0; c0 < C; c0 += kFwdChunk) {
    __syncthreads();
    for (int rr = warp; rr < kFwdRows; rr += 4) {
      const long long r = r0 + rr;
      const int c = c0 + lane;
      s_x[rr * 33 + lane] = (r < R && c < C) ? x[r * C + c] : 0.f;
    }
    __syncthreads();
    const int cn = min(kFwdChunk, C - c0);
    for (int c = 0; c < cn; ++c) {
      const float xv = s_x[r_loc * 33 + c];
      const float4* wrow = s_w4 + (size_t)(c0 + c) * KG;
#pragma unroll
      for (int j = 0; j < (kMaxKG + 1) / 2; ++j) {
        const int g = kh + 2 * j;
        if (g < KG) {
          const float4 w = wrow[g];
          acc[j].x = fmaf(xv, w.x, acc[j].x);
          acc[j].y = fmaf(xv, w.y, acc[j].y);
          acc[j].z = fmaf(xv, w.z, acc[j].z);
          acc[j].w = fmaf(xv, w.w, acc[j].w);
        }
      }
    }
  }
  const long long r = r0 + r_loc;
  if (r < R) {
#pragma unroll
    for (int j = 0; j < (kMaxKG + 1) / 2; ++j) {
      const int g = kh + 2 * j;
      if (g < KG) {
        const float v[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = g * 4 + e;
          if (k < K) out[r * K + k] = v[e] + (bias ? bias[k] : 0.f);
        }
      }
    }
  }
}

// Two small linear maps of the SAME input in one pass (AttBlock's att and cla 1x1 convolutions, models.py:137 / :141):
//   out1[r][k] = b1[k] + sum_c x[r][c] W1[k][c],   out2[r][k] = b2[k] + sum_c x[r][c] W2[k][c].
// One warp owns 32 rows, one row per lane.  The weights of both maps sit in shared memory as [c][Kp] (Kp = K1 + K2 padded
// to float4 groups): a thread reads them as warp-broadcast float4s.  x is read ONCE, coalesced (32 lanes = 32 columns
// of one row), parked in a per-warp [32][33] tile and re-read row-per-lane; the next 32-column chunk is already in
// registers while the current one is consumed, and nothing but __syncwarp() orders a warp.  (linear_small_fwd_kernel
// above staged 64 rows per block behind two __syncthreads per chunk and ran once per map: 2 x 100 us for 65 MB.)
template <int kKG>
__global__ void __launch_bounds__(128)
linear_pair_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ b1, int K1,
                       const float* __restrict__ W2, const float* __restrict__ b2, int K2, long long R, int C,
                       float* __restrict__ out1, float* __restrict__ out2) {
  extern __shared__ float4 s_lin4[];
  constexpr int Kp = kKG * 4;
  float* s_w = reinterpret_cast<float*>(s_lin4);                 // [C][Kp]: columns 0..K1-1 map 1, K1..K1+K2-1 map 2
  float* s_x = s_w + (size_t)C * Kp;                             // [4 warps][32][33]
  for (int i = threadIdx.x; i < C * Kp; i += blockDim.x) {       // coalesced rows of W1 / W2, transposed into shared memory
    const int k = i / C, c = i - k * C;
    s_w[c * Kp + k] = k < K1 ? W1[(long long)k * C + c] : (k < K1 + K2 ? W2[(long long)(k - K1) * C + c] : 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = s_x + warp * 32 * 33;
  const float4* s_w4 = reinterpret_cast<const float4*>(s_w);
  const long long n_tasks = (R + 31) / 32;
  for (long long task = (long long)blockIdx.x * 4 + warp; task < n_tasks; task += (long long)gridDim.x * 4) {
    const long long r0 = task * 32;
    float4 acc[kKG];
#pragma unroll
    for (int g = 0; g < kKG; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    float xr[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) xr[i] = (r0 + i < R && lane < C) ? __ldg(x + (r0 + i) * C + lane) : 0.f;
    for (int c0 = 0; c0 < C; c0 += 32) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 32; ++i) tile[i * 33 + lane] = xr[i];
      __syncwarp();
      if (c0 + 32 < C) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          xr[i] = (r0 + i < R && c0 + 32 + lane < C) ? __ldg(x + (r0 + i) * C + c0 + 32 + lane) : 0.f;
      }
      const int cn = min(32, C - c0);
#pragma unroll 4
      for (int c = 0; c < cn; ++c) {
        const float xv = tile[lane * 33 + c];
        const float4* wrow = s_w4 + (size_t)(c0 + c) * kKG;
#pragma unroll
        for (int g = 0; g < kKG; ++g) {
          const float4 w = wrow[g];
          acc[g].x = fmaf(xv, w.x, acc[g].x);
          acc[g].y = fmaf(xv, w.y, acc[g].y);
          acc[g].z = fmaf(xv, w.z, acc[g].z);
          acc[g].w = fmaf(xv, w.w, acc[g].w);
        }
      }
    }
    const long long r = r0 + lane;
    if (r < R) {
#pragma unroll
      for (int g = 0; g < kKG; ++g) {
        const float v[4] = {acc[g].x, acc[g].y, acc[g].z, acc[g].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = g * 4 + e;
          if (k < K1) out1[r * K1 + k] = v[e] + (b1 ? b1[k] : 0.f);
          else if (k < K1 + K2) out2[r * K2 + (k - K1)] = v[e] + (b2 ? b2[k - K1] : 0.f);
        }
      }
    }
  }
}

// dx[r][c..c+3] = sum_k dout[r][k] * W[k][c..c+3]      thread owns 4 consecutive columns, keeps its K x 4 weights in
// registers and walks the block's rows; dout rows are staged in shared memory and read as broadcasts.
template <int kKG>
__global__ void linear_small_dx_kernel(const float* __restrict__ dout, const float* __restrict__ W, long long R, int C,
                                       int K, float* __restrict__ dx, int accumulate) {
  extern __shared__ float4 s_lin4[];
  float* s_d = reinterpret_cast<float*>(s_lin4);                 // [chunk][Kp]
  constexpr int Kp = kKG * 4, chunk = 32;
  const int c = threadIdx.x * 4;
  const bool ok = c < C;
  float4 w[Kp];
#pragma unroll
  for (int k = 0; k < Kp; ++k) {                 // scalar loads: W may be a 4-byte-aligned view into a flat parameter buffer
    const float* wk = W + (long long)k * C + c;
    w[k] = (ok && k < K) ? make_float4(wk[0], wk[1], wk[2], wk[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long per = (R + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(R, r0 + per);
  for (long long rb = r0; rb < r1; rb += chunk) {
    const int n = (int)min((long long)chunk, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < n * Kp; i += blockDim.x) {
      const int j = i / Kp, k = i % Kp;
      s_d[i] = k < K ? dout[(rb + j) * K + k] : 0.f;
    }
    __syncthreads();
    if (ok)
      for (int j = 0; j < n; ++j) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* d4 = reinterpret_cast<const float4*>(s_d + j * Kp);
#pragma unroll
        for (int g = 0; g < kKG; ++g) {
          const float4 d = d4[g];
          const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            a.x = fmaf(dv[e], w[g * 4 + e].x, a.x);
            a.y = fmaf(dv[e], w[g * 4 + e].y, a.y);
            a.z = fmaf(dv[e], w[g * 4 + e].z, a.z);
            a.w = fmaf(dv[e], w[g * 4 + e].w, a.w);
          }
        }
        float4* o = reinterpret_cast<float4*>(dx + (rb + j) * C + c);
        if (accumulate) {
          const float4 p = *o;
          a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        *o = a;
      }
  }
}

// partial_w[blk][k][c..c+3] = sum over the block's rows of dout[r][k] * x[r][c..c+3];  partial_b[blk][k] = sum dout[r][k]
template <int kKG>
__global__ void linear_small_dw_kernel(const float* __restrict__ dout, const float* __restrict__ x, long long R, int C,
                                       int K, float* __restrict__ partial_w, float* __restrict__ partial_b) {
  extern __shared__ float4 s_lin4[];
  float* s_d = reinterpret_cast<float*>(s_lin4);                 // [chunk][Kp]
  constexpr int Kp = kKG * 4, chunk = 32;
  const int c = threadIdx.x * 4;
  const bool ok = c < C;
  float4 acc[Kp];
#pragma unroll
  for (int k = 0; k < Kp; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float bsum = 0.f;
  const long long per = (R + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(R, r0 + per);
  for (long long rb = r0; rb < r1; rb += chunk) {
    const int n = (int)min((long long)chunk, r1 - rb);
    __syncthreads();
    for (int i = threadIdx.x; i < n * Kp; i += blockDim.x) {
      const int j = i / Kp, k = i % Kp;
      s_d[i] = k < K ? dout[(rb + j) * K + k] : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < K)
      for (int j = 0; j < n; ++j) bsum += s_d[j * Kp + threadIdx.x];
    if (ok)
      for (int j = 0; j < n; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(x + (rb + j) * C + c);
        const float4* d4 = reinterpret_cast<const float4*>(s_d + j * Kp);
#pragma unroll
        for (int g = 0; g < kKG; ++g) {
          const float4 d = d4[g];
          const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[g * 4 + e].x = fmaf(dv[e], xv.x, acc[g * 4 + e].x);
            acc[g * 4 + e].y = fmaf(dv[e], xv.y, acc[g * 4 + e].y);
            acc[g * 4 + e].z = fmaf(dv[e], xv.z, acc[g * 4 + e].z);
            acc[g * 4 + e].w = fmaf(dv[e], xv.w, acc[g * 4 + e].w);
          }
        }
      }
  }
  if (ok)
#pragma unroll
    for (int k = 0; k < Kp; ++k)
      if (k < K) *reinterpret_cast<float4*>(partial_w + ((long long)blockIdx.x * K + k) * C + c) = acc[k];
  if (partial_b && threadIdx.x < K) partial_b[(long long)blockIdx.x * K + threadIdx.x] = bsum;
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

// AttBlock tail.  One block per clip.  exp / sigmoid of the (T, K) logits go to shared memory class-major; one warp
// per class then reduces over time with lanes striding t (fixed shuffle tree: deterministic).  Round 1 let K = 17
// threads walk their class's 125 frames twice through global memory: 90 us for 2 x 2125 values per clip.
__global__ void head_att_fwd_kernel(const float* __restrict__ att_logit, const float* __restrict__ cla_logit, int T,
                                    int K, int ratio, int sigmoid_act, float temperature,
                                    float* __restrict__ norm_att /* (B,K,T) */, float* __restrict__ cla /* (B,K,T) */,
                                    float* __restrict__ clip, float* __restrict__ frame /* (B,T*ratio,K) */) {
  extern __shared__ float s_att[];
  float* s_e = s_att;                     // [K][T]  exp(clamped logit / temperature) + 1e-6
  float* s_c = s_att + K * T;             // [K][T]  classifier activation
  const int b = blockIdx.x;
  const float* al = att_logit + (long long)b * T * K;
  const float* cl = cla_logit + (long long)b * T * K;
  float* na = norm_att + (long long)b * K * T;
  float* ca = cla + (long long)b * K * T;
  for (int i = threadIdx.x; i < T * K; i += blockDim.x) {
    const int k = i % K, t = i / K;
    const float a = fminf(fmaxf(al[i], -10.f), 10.f);
    s_e[k * T + t] = expf(a / temperature) + 1e-6f;
    s_c[k * T + t] = sigmoid_act ? sigmoidf_(cl[i]) : cl[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarps) {
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += s_e[k * T + t];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float acc = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float n = s_e[k * T + t] / s;
      const float c = s_c[k * T + t];
      na[k * T + t] = n;
      ca[k * T + t] = c;
      acc += n * c;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) clip[b * K + k] = acc;
  }
  if (frame) {
    float* fr = frame + (long long)b * T * ratio * K;
    for (int i = threadIdx.x; i < T * ratio * K; i += blockDim.x) {
      const int k = i % K, tt = i / K;
      fr[i] = s_c[k * T + tt / ratio];
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

int sed_linear_partials(void) { return sm_count() * 4; }

int sed_linear_small_fwd(const float* x, const float* W, const float* bias, long long R, int C, int K, float* out,
                         sed_stream_t stream) {
  SED_REQUIRE(x && W && out, "sed_linear_small_fwd: null pointer");
  SED_REQUIRE(K >= 1 && K <= kMaxK, "sed_linear_small_fwd: K=%d must be in [1, %d]", K, kMaxK);
  SED_REQUIRE(C >= 1 && C <= 1024, "sed_linear_small_fwd: C=%d must be in [1, 1024]", C);
  if (R == 0) return 0;
  const int Kp = (K + 3) / 4 * 4;
  const size_t smem = ((size_t)C * Kp + kFwdRows * 33) * sizeof(float);
  SED_CUDA(cudaFuncSetAttribute(linear_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_small_fwd_kernel<<<(unsigned)((R + kFwdRows - 1) / kFwdRows), 128, smem, (cudaStream_t)stream>>>(
      x, W, bias, R, C, K, out);
  SED_LAUNCH_CHECK("linear_small_fwd_kernel");
  return 0;
}

int sed_linear_pair_fwd(const float* x, const float* W1, const float* bias1, int K1, const float* W2, const float* bias2,
                        int K2, long long R, int C, float* out1, float* out2, sed_stream_t stream) {
  SED_REQUIRE(x && W1 && W2 && out1 && out2, "sed_linear_pair_fwd: null pointer");
  SED_REQUIRE(K1 >= 1 && K2 >= 1 && K1 + K2 <= kMaxPairK, "sed_linear_pair_fwd: K1 + K2 = %d must be in [2, %d]", K1 + K2,
              kMaxPairK);
  SED_REQUIRE(C >= 1 && C <= 512, "sed_linear_pair_fwd: C=%d must be in [1, 512]", C);
  if (R == 0) return 0;
  const int KG = (K1 + K2 + 3) / 4;
  const size_t smem = sizeof(float) * ((size_t)C * KG * 4 + 4 * 32 * 33);
  const long long tasks = (R + 31) / 32;
  const unsigned grid = (unsigned)((tasks + 3) / 4 < (long long)sm_count() * 2 ? (tasks + 3) / 4 : (long long)sm_count() * 2);
  cudaStream_t st = (cudaStream_t)stream;
#define SED_PAIR(G)                                                                                                  \
  case G:                                                                                                            \
    SED_CUDA(cudaFuncSetAttribute(linear_pair_fwd_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    linear_pair_fwd_kernel<G><<<grid, 128, smem, st>>>(x, W1, bias1, K1, W2, bias2, K2, R, C, out1, out2);           \
    break;
  switch (KG) {
    SED_PAIR(1) SED_PAIR(2) SED_PAIR(3) SED_PAIR(4) SED_PAIR(5) SED_PAIR(6) SED_PAIR(7) SED_PAIR(8) SED_PAIR(9) SED_PAIR(10)
    default: SED_REQUIRE(false, "sed_linear_pair_fwd: unsupported K");
  }
#undef SED_PAIR
  SED_LAUNCH_CHECK("linear_pair_fwd_kernel");
  return 0;
}

int sed_linear_small_bwd(const float* dout, const float* x, const float* W, long long R, int C, int K, float* dx,
                         int dx_accumulate, float* partial_w, float* partial_b, sed_stream_t stream) {
  SED_REQUIRE(dout && x && W, "sed_linear_small_bwd: null pointer");
  SED_REQUIRE(K >= 1 && K <= kMaxK, "sed_linear_small_bwd: K=%d must be in [1, %d]", K, kMaxK);
  SED_REQUIRE(C % 4 == 0 && C >= 4 && C <= 4096, "sed_linear_small_bwd: C=%d must be a multiple of 4, <= 4096", C);
  SED_REQUIRE(aligned(x, 16) && (!dx || aligned(dx, 16)) && (!partial_w || aligned(partial_w, 16)),
              "sed_linear_small_bwd: x, dx and partial_w must be 16-byte aligned");
  if (R == 0) return 0;
  const int threads = (C / 4 + 31) / 32 * 32;
  const int grid = (int)min(R, (long long)sed_linear_partials());
  const int kg = (K + 3) / 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
#define SED_DX(G) linear_small_dx_kernel<G><<<grid, threads, 32 * (G) * 4 * sizeof(float), st>>>(dout, W, R, C, K, dx, dx_accumulate)
    if (kg <= 2) SED_DX(2); else if (kg <= 5) SED_DX(5); else SED_DX(8);
#undef SED_DX
    SED_LAUNCH_CHECK("linear_small_dx_kernel");
  }
  if (partial_w) {
    // every one of the sed_linear_partials() partial rows must be written: blocks past R write zeros
    const int gridw = sed_linear_partials();
#define SED_DW(G) linear_small_dw_kernel<G><<<gridw, threads, 32 * (G) * 4 * sizeof(float), st>>>(dout, x, R, C, K, partial_w, partial_b)
    if (kg <= 2) SED_DW(2); else if (kg <= 5) SED_DW(5); else SED_DW(8);
#undef SED_DW
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
  const size_t smem = sizeof(float) * 2 * (size_t)T * K;
  SED_REQUIRE(smem <= 200 * 1024, "sed_head_att_fwd: T * K = %d does not fit in shared memory", T * K);
  SED_CUDA(cudaFuncSetAttribute(head_att_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_att_fwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(att_logit, cla_logit, T, K, ratio, sigmoid_act, temperature,
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
