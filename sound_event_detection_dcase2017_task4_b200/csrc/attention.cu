// attention.cu -- scaled-dot-product attention of the CNN-Transformer variants, forward and
// backward, plus the dropout+ReLU that follows the output projection.
//
// Replaces ScaledDotProductAttention.forward (/root/reference/pytorch/models.py:596-608:
// bmm, /temperature, Softmax(dim=2), Dropout(0.1), bmm) as called from MultiHead.forward
// (models.py:641-665) including its four permute().contiguous() copies: heads are addressed in
// place inside the (B*T, n_head*d) projection outputs, nothing is transposed or copied.
// The 512->512 projections around it (w_qs/w_ks/w_vs/fc) are tensor-core GEMMs (sed_gemm_tc).
//
// One CTA per (batch, head); sequence length T <= 128 (125 for 10 s clips), head dim 64, fp32.
//   forward : warp per query row: s = q.K^T/temp -> softmax -> (dropout) -> o = p.V
//   backward: pass A, warp per query row: dP = dO.V^T -> dS -> dQ; dS and the dropped-out P rows stay in shared memory;
//             pass B, warp per key row:   dK = dS^T.Q/temp, dV = Pd^T.dO
// Dropout uses a counter-based Philox4x32-10 stream keyed by (seed, element index): the mask is
// recomputed in the backward instead of being stored.  (Not bit-identical to torch's own Philox
// usage -- parity tests run with dropout disabled or compare statistics)
#include "common.cuh"

namespace sed {
namespace {

constexpr int kD = 64;            // head dimension (d_k = d_v = 64, models.py:702-707)
constexpr int kMaxT = 128;
constexpr int kAttThreads = 256;
constexpr int kWarps = kAttThreads / 32;
constexpr int kLdK = kD + 1;      // padded row stride of the K / Q tile: conflict-free column walks

__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t& hi) {
  const unsigned long long p = (unsigned long long)a * b;
  hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
// Philox4x32-10: 4 x 32 random bits for counter (idx >> 2); returns the (idx & 3)-th word.
__device__ __forceinline__ uint32_t philox_word(unsigned long long seed, unsigned long long offset,
                                                unsigned long long idx) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const unsigned long long ctr = (idx >> 2) + offset;
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5ed5ed5eu, c3 = 0;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo(0xD2511F53u, c0, hi0);
    const uint32_t lo1 = mulhilo(0xCD9E8D57u, c2, hi1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const uint32_t w = (uint32_t)(idx & 3);
  return w == 0 ? c0 : (w == 1 ? c1 : (w == 2 ? c2 : c3));
}
__device__ __forceinline__ bool keep_elem(unsigned long long seed, unsigned long long offset,
                                          unsigned long long idx, float p) {
  const float u = (float)(philox_word(seed, offset, idx) >> 8) * (1.0f / 16777216.0f);
  return u >= p;
}

struct AttParams {
  const float* q; const float* k; const float* v;   // (B*T, ld) row-major, head h at columns h*64
  int ldq, ldk, ldv;
  int B, T, H;
  float inv_temp;
  float p_drop;                                       // 0 = no dropout
  unsigned long long seed, offset;
};

// ctx (B*T, H*64) fp32; probs (B, H, T, T) fp32 softmax output BEFORE dropout (may be nullptr)
__global__ void __launch_bounds__(kAttThreads)
attention_fwd_kernel(AttParams p, float* __restrict__ ctx, float* __restrict__ probs) {
  extern __shared__ float smem[];
  float* sK = smem;                         // [T][kLdK]
  float* sV = sK + kMaxT * kLdK;            // [T][kD]
  float* sQ = sV + kMaxT * kD;              // [kWarps * 4][kD]     the warp's block of query rows
  float* sP = sQ + kWarps * 4 * kD;         // [kWarps * 4][kMaxT]  their probability rows
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < T * kD; i += kAttThreads) {
    const int t = i / kD, d = i % kD;
    const long long row = (long long)b * T + t;
    sK[t * kLdK + d] = p.k[row * p.ldk + h * kD + d];
    sV[t * kD + d] = p.v[row * p.ldv + h * kD + d];
  }
  __syncthreads();
  const float keep_scale = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  // kFB query rows per warp iteration: every K / V element read from shared memory serves kFB rows
  constexpr int kFB = 4;
  for (int i0 = warp * kFB; i0 < T; i0 += kWarps * kFB) {
#pragma unroll
    for (int r = 0; r < kFB; ++r) {
      const int i = i0 + r;
      float q0 = 0.f, q1 = 0.f;
      if (i < T) {
        const long long row = (long long)b * T + i;
        q0 = p.q[row * p.ldq + h * kD + lane];
        q1 = p.q[row * p.ldq + h * kD + lane + 32];
      }
      sQ[(warp * kFB + r) * kD + lane] = q0;
      sQ[(warp * kFB + r) * kD + lane + 32] = q1;
    }
    __syncwarp();
    float s[kFB][kMaxT / 32];
#pragma unroll
    for (int r = 0; r < kFB; ++r)
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) s[r][c] = 0.f;
    const float* qr = sQ + warp * kFB * kD;
#pragma unroll 4
    for (int d = 0; d < kD; ++d) {
      float kk[kMaxT / 32], qq[kFB];
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) kk[c] = sK[min(lane + 32 * c, T - 1) * kLdK + d];
#pragma unroll
      for (int r = 0; r < kFB; ++r) qq[r] = qr[r * kD + d];
#pragma unroll
      for (int r = 0; r < kFB; ++r)
#pragma unroll
        for (int c = 0; c < kMaxT / 32; ++c) s[r][c] = fmaf(qq[r], kk[c], s[r][c]);
    }
#pragma unroll
    for (int r = 0; r < kFB; ++r) {
      const int i = i0 + r;
      if (i >= T) break;                                    // warp-uniform
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) {
        const int j = lane + 32 * c;
        s[r][c] *= p.inv_temp;
        if (j < T) mx = fmaxf(mx, s[r][c]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) {
        const int j = lane + 32 * c;
        s[r][c] = j < T ? expf(s[r][c] - mx) : 0.f;
        sum += s[r][c];
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      const long long pbase = (((long long)b * p.H + h) * T + i) * T;
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) {
        const int j = lane + 32 * c;
        float pr = 0.f;
        if (j < T) {
          pr = s[r][c] * inv;
          if (probs) probs[pbase + j] = pr;
          if (p.p_drop > 0.f) pr = keep_elem(p.seed, p.offset, (unsigned long long)(pbase + j), p.p_drop) ? pr * keep_scale : 0.f;
        }
        sP[(warp * kFB + r) * kMaxT + j] = pr;
      }
    }
    __syncwarp();
    float o0[kFB], o1[kFB];
#pragma unroll
    for (int r = 0; r < kFB; ++r) o0[r] = o1[r] = 0.f;
    const float* pr = sP + warp * kFB * kMaxT;
    const int nr = min(kFB, T - i0);
    for (int j = 0; j < T; ++j) {
      const float va = sV[j * kD + lane], vb = sV[j * kD + lane + 32];
#pragma unroll
      for (int r = 0; r < kFB; ++r) {
        const float w = pr[min(r, nr - 1) * kMaxT + j];
        o0[r] = fmaf(w, va, o0[r]);
        o1[r] = fmaf(w, vb, o1[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kFB; ++r) {
      const int i = i0 + r;
      if (i < T) {
        const long long row = (long long)b * T + i;
        ctx[row * (p.H * kD) + h * kD + lane] = o0[r];
        ctx[row * (p.H * kD) + h * kD + lane + 32] = o1[r];
      }
    }
    __syncwarp();
  }
}

// dctx (B*T, H*64); probs (B,H,T,T) from the forward; dq/dk/dv written with the same addressing as q/k/v.
// The dS tile and the dropped-out probability tile Pd of the (batch, head) pair stay in shared memory between the
// two passes (2 x T x 128 floats), so pass B reads no global memory for them and the dropout mask is evaluated once
// per element (the first version re-read both tiles from global memory and re-ran Philox in all 32 lanes of the
// key-row warp: 2.8 ms at batch 128).
__global__ void __launch_bounds__(kAttThreads)
attention_bwd_kernel(AttParams p, const float* __restrict__ dctx, const float* __restrict__ probs,
                     float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ float smem[];
  float* sA = smem;                         // pass A: K [T][kLdK]   pass B: Q  [T][kLdK]
  float* sB = sA + kMaxT * kLdK;            // pass A: V [T][kLdK]   pass B: dO [T][kLdK]
  float* sR = sB + kMaxT * kLdK;            // [kWarps * 4][kD]   dO rows of the warp's row block
  float* sDS = sR + kWarps * 4 * kD;        // [T][kMaxT]     dS  (includes 1/temperature)
  float* sPD = sDS + kMaxT * kMaxT;         // [T][kMaxT]     softmax output after dropout
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, HD = p.H * kD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float keep_scale = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  const long long tile = ((long long)b * p.H + h) * T * T;
  for (int i = threadIdx.x; i < T * kD; i += kAttThreads) {
    const int t = i / kD, d = i % kD;
    const long long row = (long long)b * T + t;
    sA[t * kLdK + d] = p.k[row * p.ldk + h * kD + d];
    sB[t * kLdK + d] = p.v[row * p.ldv + h * kD + d];
  }
  for (int i = threadIdx.x; i < T * (kMaxT - T); i += kAttThreads) {      // key columns >= T: defined zeros
    const int r = i / (kMaxT - T), c = T + i % (kMaxT - T);
    sDS[r * kMaxT + c] = 0.f;
    sPD[r * kMaxT + c] = 0.f;
  }
  __syncthreads();
  // ---- pass A: kRB query rows per warp iteration (register blocking: every K / V element fetched from shared
  //      memory is used for kRB rows, every dO element for 4 key columns)
  constexpr int kRB = 4;
  for (int i0 = warp * kRB; i0 < T; i0 += kWarps * kRB) {
#pragma unroll
    for (int r = 0; r < kRB; ++r) {
      const int i = i0 + r;
      float v0 = 0.f, v1 = 0.f;
      if (i < T) {
        const long long row = (long long)b * T + i;
        v0 = dctx[row * HD + h * kD + lane];
        v1 = dctx[row * HD + h * kD + lane + 32];
      }
      sR[(warp * kRB + r) * kD + lane] = v0;
      sR[(warp * kRB + r) * kD + lane + 32] = v1;
    }
    __syncwarp();
    float dp[kRB][kMaxT / 32];
#pragma unroll
    for (int r = 0; r < kRB; ++r)
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) dp[r][c] = 0.f;
    const float* dr = sR + warp * kRB * kD;
#pragma unroll 4
    for (int d = 0; d < kD; ++d) {
      float vv[kMaxT / 32], dd[kRB];
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) vv[c] = sB[min(lane + 32 * c, T - 1) * kLdK + d];
#pragma unroll
      for (int r = 0; r < kRB; ++r) dd[r] = dr[r * kD + d];
#pragma unroll
      for (int r = 0; r < kRB; ++r)
#pragma unroll
        for (int c = 0; c < kMaxT / 32; ++c) dp[r][c] = fmaf(dd[r], vv[c], dp[r][c]);
    }
#pragma unroll
    for (int r = 0; r < kRB; ++r) {
      const int i = i0 + r;
      if (i >= T) break;                                    // warp-uniform
      float pr[kMaxT / 32];
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) {
        const int j = lane + 32 * c;
        float pj = 0.f, a = dp[r][c];
        if (j < T) {
          pj = probs[tile + (long long)i * T + j];
          float pd = pj;
          if (p.p_drop > 0.f) {
            const bool keep = keep_elem(p.seed, p.offset, (unsigned long long)(tile + (long long)i * T + j), p.p_drop);
            a = keep ? a * keep_scale : 0.f;
            pd = keep ? pj * keep_scale : 0.f;
          }
          sPD[i * kMaxT + j] = pd;
          dot += a * pj;
        } else {
          a = 0.f;
        }
        dp[r][c] = a;
        pr[c] = pj;
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int c = 0; c < kMaxT / 32; ++c) {
        const int j = lane + 32 * c;
        if (j < T) sDS[i * kMaxT + j] = pr[c] * (dp[r][c] - dot) * p.inv_temp;      // already includes 1/temperature
      }
    }
    __syncwarp();
    float q0[kRB], q1[kRB];
#pragma unroll
    for (int r = 0; r < kRB; ++r) q0[r] = q1[r] = 0.f;
    const int i_last = min(i0 + kRB - 1, T - 1);
    for (int j = 0; j < T; ++j) {
      const float ka = sA[j * kLdK + lane], kb = sA[j * kLdK + lane + 32];
#pragma unroll
      for (int r = 0; r < kRB; ++r) {
        const float w = sDS[min(i0 + r, i_last) * kMaxT + j];
        q0[r] = fmaf(w, ka, q0[r]);
        q1[r] = fmaf(w, kb, q1[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kRB; ++r) {
      const int i = i0 + r;
      if (i < T) {
        const long long row = (long long)b * T + i;
        dq[row * p.ldq + h * kD + lane] = q0[r];
        dq[row * p.ldq + h * kD + lane + 32] = q1[r];
      }
    }
    __syncwarp();
  }
  __syncthreads();                           // all dS / Pd rows of this tile are in shared memory
  for (int i = threadIdx.x; i < T * kD; i += kAttThreads) {
    const int t = i / kD, d = i % kD;
    const long long row = (long long)b * T + t;
    sA[t * kLdK + d] = p.q[row * p.ldq + h * kD + d];
    sB[t * kLdK + d] = dctx[row * HD + h * kD + d];
  }
  __syncthreads();
  // ---- pass B: four key rows per warp iteration   dK = dS^T . Q,  dV = Pd^T . dO
  for (int j0 = warp * 4; j0 < T; j0 += kWarps * 4) {
    float k0[4], k1[4], v0[4], v1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) k0[r] = k1[r] = v0[r] = v1[r] = 0.f;
#pragma unroll 2
    for (int i = 0; i < T; ++i) {
      const float4 ds = *reinterpret_cast<const float4*>(sDS + i * kMaxT + j0);     // warp-wide broadcasts
      const float4 pd = *reinterpret_cast<const float4*>(sPD + i * kMaxT + j0);
      const float qa = sA[i * kLdK + lane], qb = sA[i * kLdK + lane + 32];
      const float oa = sB[i * kLdK + lane], ob = sB[i * kLdK + lane + 32];
      const float dsv[4] = {ds.x, ds.y, ds.z, ds.w}, pdv[4] = {pd.x, pd.y, pd.z, pd.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        k0[r] = fmaf(dsv[r], qa, k0[r]);
        k1[r] = fmaf(dsv[r], qb, k1[r]);
        v0[r] = fmaf(pdv[r], oa, v0[r]);
        v1[r] = fmaf(pdv[r], ob, v1[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = j0 + r;
      if (j < T) {
        const long long row = (long long)b * T + j;
        dk[row * p.ldk + h * kD + lane] = k0[r];
        dk[row * p.ldk + h * kD + lane + 32] = k1[r];
        dv[row * p.ldv + h * kD + lane] = v0[r];
        dv[row * p.ldv + h * kD + lane + 32] = v1[r];
      }
    }
  }
}

// y = relu(dropout_p(x)) elementwise (models.py:664: F.relu_(self.dropout(self.fc(output))))
__global__ void dropout_relu_fwd_kernel(const float* __restrict__ x, long long n, float p, unsigned long long seed,
                                        unsigned long long offset, float* __restrict__ y) {
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (p > 0.f) v = keep_elem(seed, offset, (unsigned long long)i, p) ? v * keep_scale : 0.f;
    y[i] = fmaxf(v, 0.f);
  }
}
// dx = dy * [y > 0] / (1 - p): y > 0 implies the element was kept
__global__ void dropout_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float p,
                                        float* __restrict__ dx) {
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] * keep_scale : 0.f;
}

int check_att(const char* name, const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int B, int T,
              int H, int d) {
  SED_REQUIRE(q && k && v, "%s: null pointer", name);
  SED_REQUIRE(d == kD, "%s: head dimension %d unsupported (64 only)", name, d);
  SED_REQUIRE(T >= 1 && T <= kMaxT, "%s: sequence length %d out of range (1..%d)", name, T, kMaxT);
  SED_REQUIRE(H >= 1 && ldq >= H * d && ldk >= H * d && ldv >= H * d, "%s: bad leading dimensions", name);
  SED_REQUIRE(B >= 0, "%s: bad batch", name);
  return 0;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_attention_fwd(const float* q, const float* k, const float* v, int ldq, int ldk, int ldv, int B, int T, int H,
                      int d, float temperature, float p_drop, unsigned long long seed, unsigned long long offset,
                      float* ctx, float* probs, sed_stream_t stream) {
  if (int rc = check_att("sed_attention_fwd", q, k, v, ldq, ldk, ldv, B, T, H, d)) return rc;
  SED_REQUIRE(ctx, "sed_attention_fwd: null output");
  SED_REQUIRE(p_drop >= 0.f && p_drop < 1.f && temperature > 0.f, "sed_attention_fwd: bad scalars");
  if (B == 0) return 0;
  AttParams p{q, k, v, ldq, ldk, ldv, B, T, H, 1.0f / temperature, p_drop, seed, offset};
  const size_t smem = sizeof(float) * (size_t)(kMaxT * kLdK + kMaxT * kD + kWarps * 4 * kD + kWarps * 4 * kMaxT);
  SED_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_fwd_kernel<<<B * H, kAttThreads, smem, (cudaStream_t)stream>>>(p, ctx, probs);
  SED_LAUNCH_CHECK("attention_fwd_kernel");
  return 0;
}

int sed_attention_bwd(const float* q, const float* k, const float* v, int ldq, int ldk, int ldv, int B, int T, int H,
                      int d, float temperature, float p_drop, unsigned long long seed, unsigned long long offset,
                      const float* dctx, const float* probs, float* dq, float* dk, float* dv, sed_stream_t stream) {
  if (int rc = check_att("sed_attention_bwd", q, k, v, ldq, ldk, ldv, B, T, H, d)) return rc;
  SED_REQUIRE(dctx && probs && dq && dk && dv, "sed_attention_bwd: null pointer");
  if (B == 0) return 0;
  AttParams p{q, k, v, ldq, ldk, ldv, B, T, H, 1.0f / temperature, p_drop, seed, offset};
  const size_t smem = sizeof(float) * (size_t)(2 * kMaxT * kLdK + kWarps * 4 * kD + 2 * kMaxT * kMaxT);
  SED_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_bwd_kernel<<<B * H, kAttThreads, smem, (cudaStream_t)stream>>>(p, dctx, probs, dq, dk, dv);
  SED_LAUNCH_CHECK("attention_bwd_kernel");
  return 0;
}

int sed_dropout_relu_fwd(const float* x, long long n, float p_drop, unsigned long long seed, unsigned long long offset,
                         float* y, sed_stream_t stream) {
  SED_REQUIRE(x && y && n >= 0 && p_drop >= 0.f && p_drop < 1.f, "sed_dropout_relu_fwd: bad arguments");
  if (n == 0) return 0;
  const int grid = (int)min((n + 255) / 256, (long long)sm_count() * 16);
  dropout_relu_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, p_drop, seed, offset, y);
  SED_LAUNCH_CHECK("dropout_relu_fwd_kernel");
  return 0;
}

int sed_dropout_relu_bwd(const float* dy, const float* y, long long n, float p_drop, float* dx, sed_stream_t stream) {
  SED_REQUIRE(dy && y && dx && n >= 0 && p_drop >= 0.f && p_drop < 1.f, "sed_dropout_relu_bwd: bad arguments");
  if (n == 0) return 0;
  const int grid = (int)min((n + 255) / 256, (long long)sm_count() * 16);
  dropout_relu_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, y, n, p_drop, dx);
  SED_LAUNCH_CHECK("dropout_relu_bwd_kernel");
  return 0;
}

}  // extern "C"
