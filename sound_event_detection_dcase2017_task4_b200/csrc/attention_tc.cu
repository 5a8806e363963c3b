// attention_tc.cu -- scaled-dot-product attention of the CNN-Transformer variants on tensor cores, forward and backward.
//
// Replaces ScaledDotProductAttention.forward (/root/reference/pytorch/models.py:596-608: bmm, /temperature,
// Softmax(dim=2), Dropout(0.1), bmm) as called from MultiHead.forward (models.py:641-665), including its four
// permute().contiguous() copies: heads are addressed in place inside the (B*T, n_head*d) projection outputs.
//
// One CTA (8 warps) per (batch, head); T <= 128 (125 for 10 s clips), head dimension 64.  Every product --
// Q K^T, P V, dO V^T, dS K, dS^T Q, Pd^T dO -- is a warp-level mma.sync.m16n8k16 bf16 tile product with fp32
// accumulators.  The reference computes in fp32, so each fp32 operand is split into bf16 hi + lo and three MMAs are
// issued per tile (hi*hi + lo*hi + hi*lo; the dropped lo*lo term is 2^-16 relative): fp32-class accuracy on the bf16
// tensor pipe (tests/test_gpu_attention.py holds the kernel to 1e-4 against torch fp32).
//   * Q, K, V (and dO in the backward) are staged once per CTA in shared memory as bf16 hi / lo tiles (row stride 72
//     elements: conflict-free ldmatrix);
//   * a warp owns 16 query rows (pass A of the backward too) or 16 key rows (pass B); scores, softmax, dropout and
//     dS never leave its registers: the accumulator fragment of the first product IS the A-operand fragment of the
//     second (the flash-attention register identity), so P / dS are re-packed in place as bf16 hi / lo pairs;
//   * the backward needs dS^T and Pd^T as A operands for dK / dV.  Instead of transposing them through shared memory
//     (2 x 64 KB), pass B recomputes the transposed score-gradient tile directly as V dO^T with key rows on M -- one
//     extra tile product, no extra shared memory.
// tcgen05 is not used here on purpose: the tiles are 16 x 128 x 64 per warp with softmax between the two products,
// and a TMEM round trip per product would cost more than the MMAs (same reasoning as csrc/gru.cu).
//
// Dropout: counter-based Philox4x32-10 keyed by (seed, offset + element/4); element index = (row of the (B,H,T)
// score matrix) * 128 + key, so that the two keys a thread holds always share one Philox block.  The mask is
// recomputed in the backward, never stored.  (Statistically, not bit-wise, equivalent to torch's dropout.)
#include "common.cuh"
#include "philox.cuh"

namespace sed {
namespace {

constexpr int kD = 64;            // head dimension (d_k = d_v = 64, models.py:702-707)
constexpr int kMaxT = 128;
constexpr int kAttThreads = 256;
constexpr int kLd = 72;           // bf16 elements per shared-memory row (144 B: 8 consecutive rows hit 8 different 16 B phases)
constexpr int kTile = kMaxT * kLd;   // elements per staged tile

struct AttParams {
  const float* q; const float* k; const float* v;   // (B*T, ld) row-major, head h at columns h*64
  int ldq, ldk, ldv;
  int B, T, H;
  float inv_temp;
  float p_drop;                                       // 0 = no dropout
  unsigned long long seed, offset;
  const unsigned long long* state;                    // optional device words {seed, offset base}: see sed_b200.h
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __nv_bfloat16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (c0, c1) += (ah + al) * (bh + bl) without the lo*lo term, for the two 8-column tiles of one ldmatrix.x4 B load.  The six
// MMAs alternate between the two accumulators so that back-to-back instructions never depend on each other.
__device__ __forceinline__ void mma3x2(float (&c0)[4], float (&c1)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       const uint32_t (&bh)[4], const uint32_t (&bl)[4]) {
  mma16816(c0, ah, bh[0], bh[1]);
  mma16816(c1, ah, bh[2], bh[3]);
  mma16816(c0, al, bh[0], bh[1]);
  mma16816(c1, al, bh[2], bh[3]);
  mma16816(c0, ah, bl[0], bl[1]);
  mma16816(c1, ah, bl[2], bl[3]);
}
// fp32 pair -> bf16x2 hi and bf16x2 lo (lo = x - float(hi))
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Stage T rows x 64 columns of an fp32 matrix (row stride ld floats, first row `row0`, first column `col0`) as bf16
// hi / lo tiles [kMaxT][kLd]; rows >= T are zero.
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, long long row0, int ld, int col0, int T,
                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (int i = threadIdx.x; i < kMaxT * (kD / 4); i += kAttThreads) {
    const int t = i >> 4, c4 = (i & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < T) x = __ldg(reinterpret_cast<const float4*>(src + (row0 + t) * ld + col0 + c4));
    uint2 h, l;
    split2(x.x, x.y, h.x, l.x);
    split2(x.z, x.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + t * kLd + c4) = h;
    *reinterpret_cast<uint2*>(lo + t * kLd + c4) = l;
  }
}

// acc[16][4] (16 rows x 128 columns) = A[rows m0..m0+15][0..63] * Bm[n][0..63]^T, both staged row-major with the
// reduction index contiguous (Q K^T, dO V^T, V dO^T).  Only column tiles that start below T are computed.
__device__ __forceinline__ void tile_product_nt(float (&acc)[16][4], const __nv_bfloat16* Ah, const __nv_bfloat16* Al,
                                                const __nv_bfloat16* Bh, const __nv_bfloat16* Bl, int m0, int T, int lane) {
#pragma unroll
  for (int n = 0; n < 16; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
  const int a_off = (m0 + (lane & 15)) * kLd + (lane >> 4) * 8;
  const int b_off = ((lane & 7) + ((lane >> 4) << 3)) * kLd + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kk = 0; kk < kD / 16; ++kk) {
    uint32_t ah[4], al[4];
    ldsm_x4(ah, Ah + a_off + kk * 16);
    ldsm_x4(al, Al + a_off + kk * 16);
#pragma unroll
    for (int np = 0; np < 8; ++np) {
      if (np * 16 < T) {
        uint32_t bh[4], bl[4];
        ldsm_x4(bh, Bh + b_off + np * 16 * kLd + kk * 16);
        ldsm_x4(bl, Bl + b_off + np * 16 * kLd + kk * 16);
        mma3x2(acc[2 * np], acc[2 * np + 1], ah, al, bh, bl);
      }
    }
  }
}

// out[8][4] (16 rows x 64 columns) += sum over the 16 reduction rows k0..k0+15 of Afrag * Bm[k][0..63], Bm staged
// row-major with the OUTPUT index contiguous (P V, dS K, dS^T Q, Pd^T dO): ldmatrix.trans.
__device__ __forceinline__ void tile_accumulate_nn(float (&out)[8][4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                                   const __nv_bfloat16* Bh, const __nv_bfloat16* Bl, int k0, int lane) {
  const int b_off = (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * kLd + (lane >> 4) * 8;
#pragma unroll
  for (int nd = 0; nd < 4; ++nd) {
    uint32_t bh[4], bl[4];
    ldsm_x4_t(bh, Bh + b_off + nd * 16);
    ldsm_x4_t(bl, Bl + b_off + nd * 16);
    mma3x2(out[2 * nd], out[2 * nd + 1], ah, al, bh, bl);
  }
}

// keep flags of the two consecutive score columns (j, j+1), j even, of score row `srow` (= (b*H + h)*T + i)
__device__ __forceinline__ void keep_pair(const AttParams& p, long long srow, int j, bool& k0, bool& k1) {
  const unsigned long long idx = (unsigned long long)srow * kMaxT + j;
  const uint4 w = philox4x32_10(p.seed, (idx >> 2) + p.offset);
  const uint32_t a = (idx & 2) ? w.z : w.x, b = (idx & 2) ? w.w : w.y;
  k0 = philox_uniform(a) >= p.p_drop;
  k1 = philox_uniform(b) >= p.p_drop;
}

// ---------------------------------------------------------------------------------------------------- forward
// ctx (B*T, H*64) fp32; probs (B, H, T, T) fp32 softmax output BEFORE dropout (may be nullptr)
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_fwd_kernel(AttParams p, float* __restrict__ ctx, float* __restrict__ probs) {
  if (p.state != nullptr) { p.seed = p.state[0]; p.offset += p.state[1]; }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* sQh = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sQl = sQh + kTile;
  __nv_bfloat16* sKh = sQl + kTile;
  __nv_bfloat16* sKl = sKh + kTile;
  __nv_bfloat16* sVh = sKl + kTile;
  __nv_bfloat16* sVl = sVh + kTile;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const long long row0 = (long long)b * T;
  stage_tile(p.q, row0, p.ldq, h * kD, T, sQh, sQl);
  stage_tile(p.k, row0, p.ldk, h * kD, T, sKh, sKl);
  stage_tile(p.v, row0, p.ldv, h * kD, T, sVh, sVl);
  __syncthreads();
  const int m0 = warp * 16;
  if (m0 >= T) return;
  float s[16][4];
  tile_product_nt(s, sQh, sQl, sKh, sKl, m0, T, lane);
  // ---- softmax over the keys, rows m0 + g (elements 0, 1) and m0 + g + 8 (elements 2, 3); a row lives in 4 lanes
  const float keep_scale = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int i = m0 + g + 8 * half;
    float mx = -INFINITY;
#pragma unroll
    for (int n = 0; n < 16; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * n + 2 * t4 + e;
        float v = s[n][2 * half + e] * p.inv_temp;
        v = j < T ? v : -INFINITY;
        s[n][2 * half + e] = v;
        mx = fmaxf(mx, v);
      }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float ev = expf(s[n][2 * half + e] - mx);       // exp(-inf) = 0 for the masked columns
        s[n][2 * half + e] = ev;
        sum += ev;
      }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    const long long srow = ((long long)b * p.H + h) * T + i;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int j = 8 * n + 2 * t4;
      float p0 = s[n][2 * half] * inv, p1 = s[n][2 * half + 1] * inv;
      if (i < T && probs) {
        if (j < T) probs[srow * T + j] = p0;
        if (j + 1 < T) probs[srow * T + j + 1] = p1;
      }
      if (p.p_drop > 0.f && j < T) {
        bool k0, k1;
        keep_pair(p, srow, j, k0, k1);
        p0 = k0 ? p0 * keep_scale : 0.f;
        p1 = k1 ? p1 * keep_scale : 0.f;
      }
      s[n][2 * half] = p0;
      s[n][2 * half + 1] = p1;
    }
  }
  // ---- O = P V: the score fragments are the A fragments
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    if (ks * 16 < T) {
      uint32_t ah[4], al[4];
      split2(s[2 * ks][0], s[2 * ks][1], ah[0], al[0]);
      split2(s[2 * ks][2], s[2 * ks][3], ah[1], al[1]);
      split2(s[2 * ks + 1][0], s[2 * ks + 1][1], ah[2], al[2]);
      split2(s[2 * ks + 1][2], s[2 * ks + 1][3], ah[3], al[3]);
      tile_accumulate_nn(o, ah, al, sVh, sVl, ks * 16, lane);
    }
  }
  const int HD = p.H * kD;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int i = m0 + g + 8 * half;
    if (i < T) {
      float* dst = ctx + (row0 + i) * HD + h * kD + 2 * t4;
#pragma unroll
      for (int n = 0; n < 8; ++n) *reinterpret_cast<float2*>(dst + 8 * n) = make_float2(o[n][2 * half], o[n][2 * half + 1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- backward
// dctx (B*T, H*64); probs (B,H,T,T) from the forward; dq/dk/dv written with the same addressing as q/k/v.
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_bwd_kernel(AttParams p, const float* __restrict__ dctx, const float* __restrict__ probs,
                        float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
  if (p.state != nullptr) { p.seed = p.state[0]; p.offset += p.state[1]; }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* sQh = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sQl = sQh + kTile;
  __nv_bfloat16* sKh = sQl + kTile;
  __nv_bfloat16* sKl = sKh + kTile;
  __nv_bfloat16* sVh = sKl + kTile;
  __nv_bfloat16* sVl = sVh + kTile;
  __nv_bfloat16* sOh = sVl + kTile;          // dO
  __nv_bfloat16* sOl = sOh + kTile;
  float* sDot = reinterpret_cast<float*>(sOl + kTile);   // [kMaxT]  sum_j dP_ij P_ij of every query row
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int T = p.T, HD = p.H * kD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const long long row0 = (long long)b * T;
  const long long tile = ((long long)b * p.H + h) * T * T;
  const long long srow0 = ((long long)b * p.H + h) * T;
  const float keep_scale = p.p_drop > 0.f ? 1.0f / (1.0f - p.p_drop) : 1.0f;
  stage_tile(p.q, row0, p.ldq, h * kD, T, sQh, sQl);
  stage_tile(p.k, row0, p.ldk, h * kD, T, sKh, sKl);
  stage_tile(p.v, row0, p.ldv, h * kD, T, sVh, sVl);
  stage_tile(dctx, row0, HD, h * kD, T, sOh, sOl);
  if (threadIdx.x < kMaxT) sDot[threadIdx.x] = 0.f;
  __syncthreads();
  const int m0 = warp * 16;
  float acc[16][4];
  float out[8][4];
  // ---- pass A: the warp owns query rows m0..m0+15:  dP = dO V^T,  dS = P (dP - rowsum(dP P)) / temp,  dQ = dS K
  if (m0 < T) {
    tile_product_nt(acc, sOh, sOl, sVh, sVl, m0, T, lane);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = m0 + g + 8 * half;
      float dot = 0.f;
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int j = 8 * n + 2 * t4;
        float a0 = acc[n][2 * half], a1 = acc[n][2 * half + 1];
        float p0 = 0.f, p1 = 0.f;
        if (i < T) {
          if (j < T) p0 = __ldg(probs + tile + (long long)i * T + j);
          if (j + 1 < T) p1 = __ldg(probs + tile + (long long)i * T + j + 1);
        }
        if (p.p_drop > 0.f && j < T) {
          bool k0, k1;
          keep_pair(p, srow0 + i, j, k0, k1);
          a0 = k0 ? a0 * keep_scale : 0.f;
          a1 = k1 ? a1 * keep_scale : 0.f;
        }
        acc[n][2 * half] = a0;                      // dL/dP (before dropout), masked columns multiply p = 0 below
        acc[n][2 * half + 1] = a1;
        dot = fmaf(a0, p0, dot);
        dot = fmaf(a1, p1, dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      if (t4 == 0 && i < T) sDot[i] = dot;
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const int j = 8 * n + 2 * t4;
        float p0 = 0.f, p1 = 0.f;
        if (i < T) {
          if (j < T) p0 = __ldg(probs + tile + (long long)i * T + j);
          if (j + 1 < T) p1 = __ldg(probs + tile + (long long)i * T + j + 1);
        }
        acc[n][2 * half] = p0 * (acc[n][2 * half] - dot) * p.inv_temp;
        acc[n][2 * half + 1] = p1 * (acc[n][2 * half + 1] - dot) * p.inv_temp;
      }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) out[n][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      if (ks * 16 < T) {
        uint32_t ah[4], al[4];
        split2(acc[2 * ks][0], acc[2 * ks][1], ah[0], al[0]);
        split2(acc[2 * ks][2], acc[2 * ks][3], ah[1], al[1]);
        split2(acc[2 * ks + 1][0], acc[2 * ks + 1][1], ah[2], al[2]);
        split2(acc[2 * ks + 1][2], acc[2 * ks + 1][3], ah[3], al[3]);
        tile_accumulate_nn(out, ah, al, sKh, sKl, ks * 16, lane);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = m0 + g + 8 * half;
      if (i < T) {
        float* dst = dq + (row0 + i) * p.ldq + h * kD + 2 * t4;
#pragma unroll
        for (int n = 0; n < 8; ++n) *reinterpret_cast<float2*>(dst + 8 * n) = make_float2(out[n][2 * half], out[n][2 * half + 1]);
      }
    }
  }
  __syncthreads();                                   // every row's dot product is in sDot
  // ---- pass B: the warp owns key rows m0..m0+15; columns are query rows:  dP^T = V dO^T,
  //      dK = dS^T Q,  dV = Pd^T dO   (the transposed tiles are built in registers, 16 query rows at a time)
  if (m0 >= T) return;
  tile_product_nt(acc, sVh, sVl, sOh, sOl, m0, T, lane);
  float outv[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) { out[n][e] = 0.f; outv[n][e] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    if (ks * 16 < T) {
      float ds[2][4], pd[2][4];
#pragma unroll
      for (int nn = 0; nn < 2; ++nn) {
        const int n = 2 * ks + nn;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = m0 + g + 8 * (e >> 1);              // key (row of the transposed tile)
          const int i = 8 * n + 2 * t4 + (e & 1);           // query (column)
          float pv = 0.f, a = acc[n][e];
          bool keep = true;
          if (i < T && j < T) {
            pv = __ldg(probs + tile + (long long)i * T + j);
            if (p.p_drop > 0.f) {
              const unsigned long long idx = (unsigned long long)(srow0 + i) * kMaxT + j;
              keep = philox_uniform(philox_word(p.seed, p.offset, idx)) >= p.p_drop;
            }
          }
          const float pdrop = keep ? pv * keep_scale : 0.f;
          a = keep ? a * keep_scale : 0.f;
          const float dot = sDot[min(i, kMaxT - 1)];
          ds[nn][e] = pv * (a - dot) * p.inv_temp;
          pd[nn][e] = pdrop;
        }
      }
      uint32_t ah[4], al[4];
      split2(ds[0][0], ds[0][1], ah[0], al[0]);
      split2(ds[0][2], ds[0][3], ah[1], al[1]);
      split2(ds[1][0], ds[1][1], ah[2], al[2]);
      split2(ds[1][2], ds[1][3], ah[3], al[3]);
      tile_accumulate_nn(out, ah, al, sQh, sQl, ks * 16, lane);
      split2(pd[0][0], pd[0][1], ah[0], al[0]);
      split2(pd[0][2], pd[0][3], ah[1], al[1]);
      split2(pd[1][0], pd[1][1], ah[2], al[2]);
      split2(pd[1][2], pd[1][3], ah[3], al[3]);
      tile_accumulate_nn(outv, ah, al, sOh, sOl, ks * 16, lane);
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int j = m0 + g + 8 * half;
    if (j < T) {
      float* dstk = dk + (row0 + j) * p.ldk + h * kD + 2 * t4;
      float* dstv = dv + (row0 + j) * p.ldv + h * kD + 2 * t4;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        *reinterpret_cast<float2*>(dstk + 8 * n) = make_float2(out[n][2 * half], out[n][2 * half + 1]);
        *reinterpret_cast<float2*>(dstv + 8 * n) = make_float2(outv[n][2 * half], outv[n][2 * half + 1]);
      }
    }
  }
}

int check_att(const char* name, const void* q, const void* k, const void* v, int ldq, int ldk, int ldv, int B, int T,
              int H, int d) {
  SED_REQUIRE(q && k && v, "%s: null pointer", name);
  SED_REQUIRE(d == kD, "%s: head dimension %d unsupported (64 only)", name, d);
  SED_REQUIRE(T >= 1 && T <= kMaxT, "%s: sequence length %d out of range (1..%d)", name, T, kMaxT);
  SED_REQUIRE(H >= 1 && ldq >= H * d && ldk >= H * d && ldv >= H * d, "%s: bad leading dimensions", name);
  SED_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && aligned(q, 16) && aligned(k, 16) && aligned(v, 16),
              "%s: q/k/v must be 16-byte aligned with leading dimensions that are multiples of 4", name);
  SED_REQUIRE(B >= 0, "%s: bad batch", name);
  return 0;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_attention_fwd(const float* q, const float* k, const float* v, int ldq, int ldk, int ldv, int B, int T, int H,
                      int d, float temperature, float p_drop, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* philox_state, float* ctx, float* probs, sed_stream_t stream) {
  if (int rc = check_att("sed_attention_fwd", q, k, v, ldq, ldk, ldv, B, T, H, d)) return rc;
  SED_REQUIRE(ctx && aligned(ctx, 8), "sed_attention_fwd: null / misaligned output");
  SED_REQUIRE(p_drop >= 0.f && p_drop < 1.f && temperature > 0.f, "sed_attention_fwd: bad scalars");
  if (B == 0) return 0;
  AttParams p{q, k, v, ldq, ldk, ldv, B, T, H, 1.0f / temperature, p_drop, seed, offset, philox_state};
  const size_t smem = sizeof(__nv_bfloat16) * (size_t)(6 * kTile);
  SED_CUDA(cudaFuncSetAttribute(attention_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_fwd_kernel<<<B * H, kAttThreads, smem, (cudaStream_t)stream>>>(p, ctx, probs);
  SED_LAUNCH_CHECK("attention_tc_fwd_kernel");
  return 0;
}

int sed_attention_bwd(const float* q, const float* k, const float* v, int ldq, int ldk, int ldv, int B, int T, int H,
                      int d, float temperature, float p_drop, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* philox_state, const float* dctx, const float* probs, float* dq, float* dk,
                      float* dv, sed_stream_t stream) {
  if (int rc = check_att("sed_attention_bwd", q, k, v, ldq, ldk, ldv, B, T, H, d)) return rc;
  SED_REQUIRE(dctx && probs && dq && dk && dv, "sed_attention_bwd: null pointer");
  SED_REQUIRE(aligned(dctx, 16) && aligned(dq, 8) && aligned(dk, 8) && aligned(dv, 8), "sed_attention_bwd: misaligned pointer");
  if (B == 0) return 0;
  AttParams p{q, k, v, ldq, ldk, ldv, B, T, H, 1.0f / temperature, p_drop, seed, offset, philox_state};
  const size_t smem = sizeof(__nv_bfloat16) * (size_t)(8 * kTile) + sizeof(float) * kMaxT;
  SED_CUDA(cudaFuncSetAttribute(attention_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_bwd_kernel<<<B * H, kAttThreads, smem, (cudaStream_t)stream>>>(p, dctx, probs, dq, dk, dv);
  SED_LAUNCH_CHECK("attention_tc_bwd_kernel");
  return 0;
}

}  // extern "C"
