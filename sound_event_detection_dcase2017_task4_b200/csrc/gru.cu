// gru.cu -- bidirectional single-layer GRU recurrence (forward + BPTT) and the small layout /
// precision helpers around its tensor-core projections.
//
// Replaces cuDNN's RNN behind `nn.GRU(512, 256, num_layers=1, bias=True, batch_first=True,
// bidirectional=True)` (/root/reference/pytorch/models.py:437-438 / :529-530, call :475 / :566):
//   r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)        z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh(W_in x + b_in + r * (W_hn h + b_hn))     h' = (1 - z) * n + z * h        (gate order r, z, n)
//
// Split of the work:
//   * input projections for all time steps and both directions: ONE tensor-core GEMM
//     (B*T, 512) x (512, 1536) with the 3-way bf16 operand split (sed_split_bf16x3 + sed_gemm_tc);
//   * the recurrence: per time step ONE launch covering both directions; CTA = (direction,
//     16 batch rows, 32 hidden units); its 96 x 256 slice of W_hh and the 16 previous hidden rows sit
//     in shared memory; gate math in fp32.  Saves (r, z, n, W_hn h + b_hn) for the backward.
//   * BPTT: per step ONE launch; CTA = (direction, 16 batch rows, 32 columns of W_hh); the gate
//     gradients of the 16 rows are rebuilt in shared memory, then dh_prev = dgh . W_hh (+ dh * z).
//   * weight gradients after the loop: tensor-core TN GEMMs (sed_gemm_tn_tc) over all (b, t).
#include "common.cuh"

namespace sed {
namespace {

constexpr int kBT = 16;       // batch rows per CTA
constexpr int kHT = 32;       // hidden units (fwd) / W_hh columns (bwd) per CTA
constexpr int kGruThreads = kBT * kHT;

__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }

// Gx (B,T,2,3H) | Whh (2,3H,H) | bhh (2,3H) | out (B,T,2H) | gates (B,T,2,4,H) = r,z,n,ghn
__global__ void __launch_bounds__(kGruThreads)
gru_fwd_step_kernel(const float* __restrict__ Gx, const float* __restrict__ Whh, const float* __restrict__ bhh,
                    float* __restrict__ out, float* __restrict__ gates, int B, int T, int H, int s) {
  extern __shared__ __align__(16) float smem[];
  const int ldw = H + 4;
  float* sW = smem;                      // [3*kHT][H+4]
  float* sH = smem + 3 * kHT * ldw;      // [kBT][H]
  const int d = blockIdx.z, b0 = blockIdx.y * kBT, j0 = blockIdx.x * kHT;
  const int tt = d == 0 ? s : T - 1 - s;
  const int tp = d == 0 ? tt - 1 : tt + 1;
  const int tid = threadIdx.x;
  const float* W = Whh + (long long)d * 3 * H * H;
  for (int i = tid; i < 3 * kHT * (H / 4); i += kGruThreads) {
    const int row = i / (H / 4), k4 = i % (H / 4);
    const int g = row / kHT, jl = row % kHT;
    const float4 v = *reinterpret_cast<const float4*>(W + ((long long)(g * H + j0 + jl)) * H + k4 * 4);
    *reinterpret_cast<float4*>(sW + row * ldw + k4 * 4) = v;
  }
  for (int i = tid; i < kBT * (H / 4); i += kGruThreads) {
    const int bl = i / (H / 4), k4 = i % (H / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s > 0 && b0 + bl < B)
      v = *reinterpret_cast<const float4*>(out + ((long long)(b0 + bl) * T + tp) * 2 * H + d * H + k4 * 4);
    *reinterpret_cast<float4*>(sH + bl * H + k4 * 4) = v;
  }
  __syncthreads();
  const int bl = tid / kHT, jl = tid % kHT;
  const int b = b0 + bl, j = j0 + jl;
  if (b >= B) return;
  float ar = 0.f, az = 0.f, an = 0.f;
  const float* hrow = sH + bl * H;
  const float* wr = sW + (0 * kHT + jl) * ldw;
  const float* wz = sW + (1 * kHT + jl) * ldw;
  const float* wn = sW + (2 * kHT + jl) * ldw;
  for (int k = 0; k < H; k += 4) {
    const float4 h4 = *reinterpret_cast<const float4*>(hrow + k);
    const float4 r4 = *reinterpret_cast<const float4*>(wr + k);
    const float4 z4 = *reinterpret_cast<const float4*>(wz + k);
    const float4 n4 = *reinterpret_cast<const float4*>(wn + k);
    ar = fmaf(h4.x, r4.x, ar); ar = fmaf(h4.y, r4.y, ar); ar = fmaf(h4.z, r4.z, ar); ar = fmaf(h4.w, r4.w, ar);
    az = fmaf(h4.x, z4.x, az); az = fmaf(h4.y, z4.y, az); az = fmaf(h4.z, z4.z, az); az = fmaf(h4.w, z4.w, az);
    an = fmaf(h4.x, n4.x, an); an = fmaf(h4.y, n4.y, an); an = fmaf(h4.z, n4.z, an); an = fmaf(h4.w, n4.w, an);
  }
  const float* bh = bhh + d * 3 * H;
  const float* gx = Gx + (((long long)b * T + tt) * 2 + d) * 3 * H;
  const float ghn = an + bh[2 * H + j];
  const float r = sigm(gx[j] + ar + bh[j]);
  const float z = sigm(gx[H + j] + az + bh[H + j]);
  const float n = tanhf(gx[2 * H + j] + r * ghn);
  const float hp = hrow[j];
  out[((long long)b * T + tt) * 2 * H + d * H + j] = (1.f - z) * n + z * hp;
  float* gs = gates + (((long long)b * T + tt) * 2 + d) * 4 * H;
  gs[j] = r; gs[H + j] = z; gs[2 * H + j] = n; gs[3 * H + j] = ghn;
}

// dout (B,T,2H) | carry_in/out (2,B,H) | dGx, dGh (B,T,2,3H) | Hprev (B,T,2,H)
__global__ void __launch_bounds__(kGruThreads)
gru_bwd_step_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ gates,
                    const float* __restrict__ Whh, const float* __restrict__ carry_in, float* __restrict__ carry_out,
                    float* __restrict__ dGx, float* __restrict__ dGh, float* __restrict__ Hprev, int B, int T, int H,
                    int bs) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                      // [3H][kHT]
  float* sD = smem + 3 * H * kHT;        // [kBT][3H]   dgh
  float* sC = sD + kBT * 3 * H;          // [kBT][kHT]  dh*z for this CTA's columns
  const int d = blockIdx.z, b0 = blockIdx.y * kBT, k0 = blockIdx.x * kHT;
  const int tt = d == 0 ? T - 1 - bs : bs;
  const int tp = d == 0 ? tt - 1 : tt + 1;             // time index of h_prev (may be out of range)
  const bool has_prev = tp >= 0 && tp < T;
  const int tid = threadIdx.x;
  const float* W = Whh + (long long)d * 3 * H * H;
  for (int i = tid; i < 3 * H * (kHT / 4); i += kGruThreads) {
    const int row = i / (kHT / 4), c4 = i % (kHT / 4);
    *reinterpret_cast<float4*>(sW + row * kHT + c4 * 4) =
        *reinterpret_cast<const float4*>(W + (long long)row * H + k0 + c4 * 4);
  }
  for (int i = tid; i < kBT * H; i += kGruThreads) {
    const int bl = i / H, j = i % H;
    const int b = b0 + bl;
    float dr_pre = 0.f, dz_pre = 0.f, dn_pre = 0.f, r = 0.f, dhz = 0.f, hp = 0.f;
    if (b < B) {
      const long long bt = (long long)b * T + tt;
      const float* gs = gates + (bt * 2 + d) * 4 * H;
      r = gs[j];
      const float z = gs[H + j], n = gs[2 * H + j], ghn = gs[3 * H + j];
      hp = has_prev ? out[((long long)b * T + tp) * 2 * H + d * H + j] : 0.f;
      float dh = dout[bt * 2 * H + d * H + j];
      if (bs > 0) dh += carry_in[((long long)d * B + b) * H + j];
      dn_pre = dh * (1.f - z) * (1.f - n * n);
      dz_pre = dh * (hp - n) * z * (1.f - z);
      dr_pre = dn_pre * ghn * r * (1.f - r);
      dhz = dh * z;
      if (blockIdx.x == 0) {
        float* gx = dGx + (bt * 2 + d) * 3 * H;
        float* gh = dGh + (bt * 2 + d) * 3 * H;
        gx[j] = dr_pre; gx[H + j] = dz_pre; gx[2 * H + j] = dn_pre;
        gh[j] = dr_pre; gh[H + j] = dz_pre; gh[2 * H + j] = dn_pre * r;
        Hprev[(bt * 2 + d) * H + j] = hp;
      }
    }
    sD[bl * 3 * H + j] = dr_pre;
    sD[bl * 3 * H + H + j] = dz_pre;
    sD[bl * 3 * H + 2 * H + j] = dn_pre * r;
    if (j >= k0 && j < k0 + kHT) sC[bl * kHT + (j - k0)] = dhz;
  }
  __syncthreads();
  const int bl = tid / kHT, kl = tid % kHT;
  const int b = b0 + bl;
  if (b >= B) return;
  float acc = 0.f;
  const float* drow = sD + bl * 3 * H;
  for (int row = 0; row < 3 * H; row += 4) {
    const float4 d4 = *reinterpret_cast<const float4*>(drow + row);
    acc = fmaf(d4.x, sW[(row + 0) * kHT + kl], acc);
    acc = fmaf(d4.y, sW[(row + 1) * kHT + kl], acc);
    acc = fmaf(d4.z, sW[(row + 2) * kHT + kl], acc);
    acc = fmaf(d4.w, sW[(row + 3) * kHT + kl], acc);
  }
  carry_out[((long long)d * B + b) * H + k0 + kl] = acc + sC[bl * kHT + kl];
}

// ================================================================ persistent kernels (H = 256)
// The per-step kernels above re-read their W_hh slice from L2 at every time step (98 KB x 256 CTAs
// x 125 steps) and pay a launch per step.  For the reference configuration (hidden size 256) the whole
// recurrence instead runs in ONE cooperative launch: a group of 8 CTAs owns (32 batch rows, one direction);
// CTA r of the group keeps the slice of W_hh for hidden units [32r, 32r+32) in shared memory for all T
// steps.  (A thread-block-cluster version with DSMEM exchange was measured first: only 8 clusters of 8 CTAs x 164 KB
// are co-resident on a B200, which leaves more than half of the SMs idle -- profiles/r01_gru_cluster.md -- so the
// exchange goes through L2.)
//
// Per-step exchange (round 2): flagged 8-byte words, no barrier.  Round 1 met at a counter in global memory (bar.sync,
// __threadfence, atomicAdd, acquire-poll, bar.sync) and then read the step's rows back from the output tensor: ncu
// (profiles/r02_gru_ncu.md) put 39 % of all warp time into that meeting -- 1.25 us per step in the fence alone, which
// has to wait for the 10 KB of gate / output stores in front of it, 0.6 us in the L1-invalidating acquire polls.  Now
// every produced value travels as {bf16 hi | bf16 lo, tag} in one 8-byte word of a double-buffered exchange area
// (tag = step + 1, area zeroed by the host entry point): a consumer polls the words it needs until their tags match
// and has data and "ready" in the same L2 round trip -- no fence, no atomic, and the hi/lo split the tensor-core
// product needs is made once by the producer instead of by all 8 consumers.  The 16-byte vector accesses used here
// carry two such 64-bit words (each single-copy atomic) and both tags are checked.
// Slot reuse: slot (s & 1) is rewritten at step s + 2, which a producer only reaches after it has consumed step s + 1
// from all 8 CTAs, and each of those wrote its step s + 1 values after all of its threads had finished reading step s.
// The two 16-row halves of a CTA own disjoint batch rows, i.e. independent recurrences: while one half polls, the
// other computes.
constexpr int kPH = 256;           // hidden size handled by the persistent kernels
constexpr int kPB = 32;            // batch rows per group
constexpr int kPJ = 32;            // hidden units per CTA
constexpr int kPGroup = kPH / kPJ; // 8 CTAs per group
constexpr int kPThreads = 256;

__device__ __forceinline__ void half_barrier_sync(int bar_id) {
  asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
}
// 16-byte exchange accesses as TWO 64-bit elements {value | tag << 32}: a naturally aligned 64-bit scalar element of a
// vector access is single-copy atomic in the PTX memory model, so a value and its tag are always observed together.
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
  unsigned long long a, b;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  return make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
}
__device__ __forceinline__ void st_relaxed_v4(uint4* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  const unsigned long long a = (unsigned long long)x | ((unsigned long long)y << 32);
  const unsigned long long b = (unsigned long long)z | ((unsigned long long)w << 32);
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// fp32 -> {bf16 hi (low half-word), bf16 lo = bf16(v - hi) (high half-word)}: the two terms of the operand split
__device__ __forceinline__ uint32_t pack_hilo(float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

struct GruPersist {
  uint4* xchg;              // [tiles][2 directions][2 slots][32 rows][values / 2] exchange area (zeroed by the host entry point)
  int B, T, b_base;
};

// ---- warp-level tensor-core pieces of the persistent kernels ----------------------------------------------------
// The recurrent products are tiny per CTA and per step (32 x 96 x 256 forward, 32 x 32 x 768 backward) and sit on a
// 125-step dependency chain, so they run as warp-level mma.sync.m16n8k16 tiles straight out of shared memory (a
// tcgen05 tile would need M >= 64 rows and a TMEM round trip per step).  fp32 operands are split into bf16 hi + lo
// parts and three products are accumulated (hi*hi + lo*hi + hi*lo): fp32-class accuracy on the bf16 pipe.
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
  hi = pack_bf16x2(h0, h1);
  lo = pack_bf16x2(v0 - h0, v1 - h1);
}
// fp32 row-major [rows][cols] (global, read through L2) -> bf16 hi / lo shared tiles with row stride `ld` elements
__device__ __forceinline__ void stage_split(const float* __restrict__ src, long long src_stride, int rows, int cols,
                                            bool have, int valid_rows, __nv_bfloat16* __restrict__ s_hi,
                                            __nv_bfloat16* __restrict__ s_lo, int ld, int tid, int nthreads) {
  const int c4n = cols >> 2;
  for (int i = tid; i < rows * c4n; i += nthreads) {
    const int r = i / c4n, c4 = i - r * c4n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have && r < valid_rows) v = __ldcg(reinterpret_cast<const float4*>(src + r * src_stride + c4 * 4));
    uint32_t h0, l0, h1, l1;
    split2(v.x, v.y, h0, l0);
    split2(v.z, v.w, h1, l1);
    *reinterpret_cast<uint2*>(s_hi + r * ld + c4 * 4) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(s_lo + r * ld + c4 * 4) = make_uint2(l0, l1);
  }
}
// The per-step exchange tile: 16 rows of kV4PerRow 16-byte vectors {hi|lo, tag, hi|lo, tag} in an exchange slot (L2).
// A thread issues its loads in batches of 8 independent requests, checks the 16 tags, and repeats the batch until all
// match; then the hi / lo halves go to the shared operand tiles.  Rows >= valid_rows (ragged last batch tile) are
// never produced: they are zero-filled without polling.
template <int kV4PerRow>
__device__ __forceinline__ void stage_flagged_tile(const uint4* __restrict__ slot, uint32_t tag, bool have,
                                                   int valid_rows, __nv_bfloat16* __restrict__ s_hi,
                                                   __nv_bfloat16* __restrict__ s_lo, int ld, int ht) {
  constexpr int kPer = 16 * kV4PerRow / 128;          // vectors per thread; vector u of thread ht: index u * 128 + ht
  constexpr int kPerRow = kV4PerRow / 128;            // ... which lies in row u / kPerRow (ht < 128)
  constexpr int kBatch = 16;                          // requests in flight per thread (8: 0.03 ms slower per launch)
  static_assert(kPer % kBatch == 0 && kV4PerRow % 128 == 0, "tile must split into whole batches per thread");
  const int n_valid = have ? min(kPer, max(valid_rows, 0) * kPerRow) : 0;
#pragma unroll 1
  for (int base = 0; base < kPer; base += kBatch) {
    uint4 v[kBatch];
    if (base < n_valid) {
      unsigned int spins = 0;
      bool ok;
      do {
        // all requests of the batch first, THEN the tag checks: a check right behind its load stalls the in-order
        // issue of the next load on a full L2 round trip (measured: it serialized every load of the step)
#pragma unroll
        for (int u = 0; u < kBatch; ++u) v[u] = ld_relaxed_v4(slot + min(base + u, n_valid - 1) * 128 + ht);
        unsigned int bad = 0u;
#pragma unroll
        for (int u = 0; u < kBatch; ++u) bad |= (v[u].y ^ tag) | (v[u].w ^ tag);
        ok = bad == 0u;
        if (!ok && ++spins > (1u << 24)) { printf("sed: GRU exchange timed out (tag %u)\n", tag); __trap(); }
      } while (!ok);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = (base + u) * 128 + ht;
      const int r = i / kV4PerRow, c4 = i - r * kV4PerRow;
      uint32_t hi = 0u, lo = 0u;
      if (base + u < n_valid) {
        hi = __byte_perm(v[u].x, v[u].z, 0x5410);
        lo = __byte_perm(v[u].x, v[u].z, 0x7632);
      }
      *reinterpret_cast<uint32_t*>(s_hi + r * ld + c4 * 2) = hi;
      *reinterpret_cast<uint32_t*>(s_lo + r * ld + c4 * 2) = lo;
    }
  }
}
constexpr int kLdK = kPH + 8;          // bf16 row stride of the K = 256 tiles (528 B: conflict-free fragment loads)
constexpr int kLdB = 3 * kPJ + 8;      // bf16 row stride of the K = 96 tiles of the backward (208 B: conflict-free)
constexpr int kBwdXchgValues = kPGroup * kPH;   // backward exchange, per batch row: 8 producers x 256 partial sums

// Forward.  CTA r of a group owns hidden units [32r, 32r+32) of 32 batch rows: a 32 x 96 x 256 product per step
// (gate columns ordered r | z | n).  Warp w: batch rows 16*(w&1).., unit octet w>>1, i.e. the three n-tiles
// {o, 4+o, 8+o}: every thread ends up with r, z, n pre-activations of the SAME (row, unit) pairs, does the gate math
// in registers and keeps its four h values for the next step's z*h term.
__global__ void __launch_bounds__(kPThreads, 1)
gru_fwd_persistent_kernel(const float* __restrict__ Gx, const float* __restrict__ Whh, const float* __restrict__ bhh,
                          float* out, float* __restrict__ gates, GruPersist q) {
  extern __shared__ __align__(16) uint8_t smem_gru[];
  __nv_bfloat16* sWh = reinterpret_cast<__nv_bfloat16*>(smem_gru);      // [96][kLdK]  W_hh slice, hi
  __nv_bfloat16* sWl = sWh + 3 * kPJ * kLdK;                            //             lo
  __nv_bfloat16* sHh = sWl + 3 * kPJ * kLdK;                            // [32][kLdK]  h_{t-1}, hi
  __nv_bfloat16* sHl = sHh + kPB * kLdK;                                //             lo
  constexpr int H = kPH;
  const int B = q.B, T = q.T;
  const int j0 = blockIdx.x * kPJ, b0 = q.b_base + blockIdx.y * kPB, d = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const float* W = Whh + (long long)d * 3 * H * H;
  for (int g = 0; g < 3; ++g)          // rows g*32 .. g*32+31 of the slice = W_hh rows g*H + j0 ..
    stage_split(W + ((long long)g * H + j0) * H, H, kPJ, H, true, kPJ, sWh + g * kPJ * kLdK, sWl + g * kPJ * kLdK, kLdK,
                tid, kPThreads);
  const int mt = warp & 1, oct = warp >> 1;                             // half (16 batch rows) / unit octet
  const int ht = oct * 32 + lane;                                       // thread index within the half (0..127)
  uint4* xbase = q.xchg + (long long)(blockIdx.y * 2 + d) * 2 * kPB * (kPH / 2);   // this group's two slots
  __syncthreads();                                                      // the W_hh slice is staged
  const int row0 = mt * 16 + gid;                                       // this thread's rows: row0, row0 + 8
  const int jl = oct * 8 + tig * 2;                                     // its units: j0 + jl, j0 + jl + 1
  const int j = j0 + jl;
  const float* bh = bhh + d * 3 * H;
  // (scalar loads: a bias may be a 4-byte-aligned view into a flat parameter buffer)
  const float2 bh_r = make_float2(bh[j], bh[j + 1]), bh_z = make_float2(bh[H + j], bh[H + j + 1]),
               bh_n = make_float2(bh[2 * H + j], bh[2 * H + j + 1]);
  float hprev[2][2] = {{0.f, 0.f}, {0.f, 0.f}};                         // [row][unit]
  for (int s = 0; s < T; ++s) {
    const int tt = d == 0 ? s : T - 1 - s;
    // this step's input projections: issued now, consumed after the recurrent product
    float2 gx[2][3];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int b = b0 + row0 + rr * 8;
#pragma unroll
      for (int g = 0; g < 3; ++g) gx[rr][g] = make_float2(0.f, 0.f);
      if (b < B) {
        const float* gp = Gx + (((long long)b * T + tt) * 2 + d) * 3 * H + j;
#pragma unroll
        for (int g = 0; g < 3; ++g) gx[rr][g] = __ldg(reinterpret_cast<const float2*>(gp + g * H));
      }
    }
    // h_{t-1} of this half's 16 rows, produced by all 8 CTAs of the group: slot (s - 1) & 1, tag s
    half_barrier_sync(1 + mt);                                          // last step's reads of the operand tile are done
    stage_flagged_tile<kPH / 2>(xbase + ((s + 1) & 1) * kPB * (kPH / 2) + mt * 16 * (kPH / 2), (uint32_t)s, s > 0,
                                B - b0 - mt * 16, sHh + mt * 16 * kLdK, sHl + mt * 16 * kLdK, kLdK, ht);
    half_barrier_sync(1 + mt);
    float acc[3][4];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[g][i] = 0.f;
    const __nv_bfloat16* ah = sHh + row0 * kLdK + tig * 2;
    const __nv_bfloat16* al = sHl + row0 * kLdK + tig * 2;
    const int brow = oct * 8 + gid;                                     // B fragment row (n index) within a gate block
#pragma unroll 4
    for (int k0 = 0; k0 < H; k0 += 16) {
      uint32_t a_hi[4], a_lo[4];
      a_hi[0] = *reinterpret_cast<const uint32_t*>(ah + k0);
      a_hi[1] = *reinterpret_cast<const uint32_t*>(ah + 8 * kLdK + k0);
      a_hi[2] = *reinterpret_cast<const uint32_t*>(ah + k0 + 8);
      a_hi[3] = *reinterpret_cast<const uint32_t*>(ah + 8 * kLdK + k0 + 8);
      a_lo[0] = *reinterpret_cast<const uint32_t*>(al + k0);
      a_lo[1] = *reinterpret_cast<const uint32_t*>(al + 8 * kLdK + k0);
      a_lo[2] = *reinterpret_cast<const uint32_t*>(al + k0 + 8);
      a_lo[3] = *reinterpret_cast<const uint32_t*>(al + 8 * kLdK + k0 + 8);
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const __nv_bfloat16* wh = sWh + (g * kPJ + brow) * kLdK + k0 + tig * 2;
        const __nv_bfloat16* wl = sWl + (g * kPJ + brow) * kLdK + k0 + tig * 2;
        uint32_t b_hi[2], b_lo[2];
        b_hi[0] = *reinterpret_cast<const uint32_t*>(wh);
        b_hi[1] = *reinterpret_cast<const uint32_t*>(wh + 8);
        b_lo[0] = *reinterpret_cast<const uint32_t*>(wl);
        b_lo[1] = *reinterpret_cast<const uint32_t*>(wl + 8);
        mma_bf16_m16n8k16(acc[g], a_hi, b_hi);
        mma_bf16_m16n8k16(acc[g], a_lo, b_hi);
        mma_bf16_m16n8k16(acc[g], a_hi, b_lo);
      }
    }
    // accumulator element i of a thread: row = row0 + 8*(i>>1), unit = j + (i&1)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int b = b0 + row0 + rr * 8;
      if (b < B) {
        float hn[2], rv[2], zv[2], nv[2], gv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = rr * 2 + u;
          const float ghn = acc[2][i] + (u ? bh_n.y : bh_n.x);
          const float r = sigm((u ? gx[rr][0].y : gx[rr][0].x) + acc[0][i] + (u ? bh_r.y : bh_r.x));
          const float z = sigm((u ? gx[rr][1].y : gx[rr][1].x) + acc[1][i] + (u ? bh_z.y : bh_z.x));
          const float n = tanhf((u ? gx[rr][2].y : gx[rr][2].x) + r * ghn);
          hn[u] = (1.f - z) * n + z * hprev[rr][u];
          rv[u] = r; zv[u] = z; nv[u] = n; gv[u] = ghn;
          hprev[rr][u] = hn[u];
        }
        // first the exchange word the other 7 CTAs (and this one) are polling for, then the tensors of record
        st_relaxed_v4(xbase + (s & 1) * kPB * (kPH / 2) + (row0 + rr * 8) * (kPH / 2) + (j >> 1), pack_hilo(hn[0]),
                       (uint32_t)(s + 1), pack_hilo(hn[1]), (uint32_t)(s + 1));
        *reinterpret_cast<float2*>(out + ((long long)b * T + tt) * 2 * H + d * H + j) = make_float2(hn[0], hn[1]);
        float* gs = gates + (((long long)b * T + tt) * 2 + d) * 4 * H + j;
        *reinterpret_cast<float2*>(gs) = make_float2(rv[0], rv[1]);
        *reinterpret_cast<float2*>(gs + H) = make_float2(zv[0], zv[1]);
        *reinterpret_cast<float2*>(gs + 2 * H) = make_float2(nv[0], nv[1]);
        *reinterpret_cast<float2*>(gs + 3 * H) = make_float2(gv[0], gv[1]);
      }
    }
    // (the next iteration's leading half-barrier orders this step's shared-memory reads before the refill)
  }
}

// BPTT in one cooperative launch, same grouping: CTA r owns hidden units [32r, 32r+32) of 32 batch rows.
//   dh_prev[b, k] = sum over (gate g, unit j) of dGh[b, g, j] * W_hh[g H + j, k]  +  dh[b, k] * z[b, k]
// Round 1 gave CTA r the 32 COLUMNS k of that product, so it needed the group's full dGh rows every step: 16 rows x
// 768 values per half, 96 KB as flagged words.  Here CTA r contracts over ITS OWN 96 (gate, unit) rows -- the gate
// gradients it has just computed, straight from shared memory -- against W_hh[rows of its units, all 256 columns] and
// gets a PARTIAL sum for every column; the exchange carries the partial sums to the CTA that owns the column
// (16 rows x 32 columns x 8 producers per half = 32 KB, one batch of 16 requests per thread), which adds the eight
// partials in producer order (deterministic) and the dh * z term of its own units.  The product of a step no longer
// waits for the exchange, only the (cheap) gate math of the next step does.
// Exchange slot layout: [slot][half][consumer CTA][producer CTA][16 rows][16 column pairs] x {p0, tag, p1, tag}, tag =
// step + 1; slot reuse is safe for the reason given at the top of this section.
__global__ void __launch_bounds__(kPThreads, 1)
gru_bwd_persistent_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ gates,
                          const float* __restrict__ Whh, float* __restrict__ dGx, float* dGh,
                          float* __restrict__ Hprev, __nv_bfloat16* __restrict__ dGx16, __nv_bfloat16* __restrict__ dGh16,
                          __nv_bfloat16* __restrict__ Hprev16, float* __restrict__ bias_partial, GruPersist q) {
  extern __shared__ __align__(16) uint8_t smem_gru[];
  __nv_bfloat16* sWh = reinterpret_cast<__nv_bfloat16*>(smem_gru);      // [256][kLdB]  W_hh[g H + u0 + jl][n] at [n][g 32 + jl], hi
  __nv_bfloat16* sWl = sWh + kPH * kLdB;                                //              lo
  __nv_bfloat16* sAh = sWl + kPH * kLdB;                                // [32][kLdB]   this step's dGh of the own units, hi
  __nv_bfloat16* sAl = sAh + kPB * kLdB;                                //              lo
  constexpr int H = kPH;
  const int B = q.B, T = q.T;
  const int r_cta = blockIdx.x, u0 = r_cta * kPJ, b0 = q.b_base + blockIdx.y * kPB, d = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const float* W = Whh + (long long)d * 3 * H * H;
  for (int i = tid; i < 3 * kPJ * (H / 4); i += kPThreads) {            // coalesced rows of W_hh, transposed into sW
    const int row = i / (H / 4), n4 = i - row * (H / 4);                // row = g * 32 + jl
    const int g = row / kPJ, jl = row - g * kPJ;
    const float4 v = *reinterpret_cast<const float4*>(W + ((long long)g * H + u0 + jl) * H + n4 * 4);
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(vv[e]);
      sWh[(n4 * 4 + e) * kLdB + row] = hi;
      sWl[(n4 * 4 + e) * kLdB + row] = __float2bfloat16_rn(vv[e] - __bfloat162float(hi));
    }
  }
  const int mt = warp & 1, oct = warp >> 1;
  // exchange area of this group: [2 slots][2 halves][8 consumers][8 producers][16 rows][16 pairs]
  constexpr int kPairVecs = 16 * 16;                                    // vectors of one (consumer, producer) block
  uint4* xbase = q.xchg + (long long)(blockIdx.y * 2 + d) * 2 * 2 * kPGroup * kPGroup * kPairVecs;
  __syncthreads();                                                      // the W_hh slice is staged
  const int row0 = mt * 16 + gid;
  const int jl = oct * 8 + tig * 2;                                     // this thread's units: u0 + jl, u0 + jl + 1
  const int j = u0 + jl;
  float dhz[2][2] = {{0.f, 0.f}, {0.f, 0.f}};                           // dh * z of the previous step (own units)
  float bsum[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};  // bias gradients: sum over (rows, t) of dr, dz, dn, dn*r
  for (int bs = 0; bs < T; ++bs) {
    const int tt = d == 0 ? T - 1 - bs : bs;
    const int tp = d == 0 ? tt - 1 : tt + 1;
    const bool has_prev = tp >= 0 && tp < T;
    // ---- this step's saved activations: requested before the exchange is polled
    float2 r2[2], z2[2], n2[2], g2[2], hp2[2], do2[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int b = b0 + row0 + rr * 8;
      r2[rr] = z2[rr] = n2[rr] = g2[rr] = hp2[rr] = do2[rr] = make_float2(0.f, 0.f);
      if (b < B) {
        const long long bt = (long long)b * T + tt;
        const float* gs = gates + (bt * 2 + d) * 4 * H + j;
        r2[rr] = __ldg(reinterpret_cast<const float2*>(gs));
        z2[rr] = __ldg(reinterpret_cast<const float2*>(gs + H));
        n2[rr] = __ldg(reinterpret_cast<const float2*>(gs + 2 * H));
        g2[rr] = __ldg(reinterpret_cast<const float2*>(gs + 3 * H));
        if (has_prev) hp2[rr] = __ldg(reinterpret_cast<const float2*>(out + ((long long)b * T + tp) * 2 * H + d * H + j));
        do2[rr] = __ldg(reinterpret_cast<const float2*>(dout + bt * 2 * H + d * H + j));
      }
    }
    // ---- dh flowing back from the later time step: eight partial sums per (row, unit) + dh * z
    float carry[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (bs > 0) {
      const uint4* xs = xbase + ((((bs - 1) & 1) * 2 + mt) * kPGroup + r_cta) * kPGroup * kPairVecs + gid * 16 + (jl >> 1);
      const uint32_t tag = (uint32_t)bs;
      const bool need[2] = {b0 + row0 < B, b0 + row0 + 8 < B};
      uint4 v[16];                                                      // [producer][rr]
      unsigned int spins = 0;
      bool ok;
      do {
#pragma unroll
        for (int rp = 0; rp < kPGroup; ++rp)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr)
            v[rp * 2 + rr] = ld_relaxed_v4(xs + rp * kPairVecs + (need[rr] ? rr * 8 * 16 : 0));
        unsigned int bad = 0u;
#pragma unroll
        for (int u = 0; u < 16; ++u) bad |= (v[u].y ^ tag) | (v[u].w ^ tag);
        ok = bad == 0u || !need[0];                                     // (rows come in order: !need[0] => no row of this thread)
        if (!ok && ++spins > (1u << 24)) { printf("sed: GRU exchange timed out (tag %u)\n", tag); __trap(); }
      } while (!ok);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int rp = 0; rp < kPGroup; ++rp) {
          sx += __uint_as_float(v[rp * 2 + rr].x);
          sy += __uint_as_float(v[rp * 2 + rr].z);
        }
        carry[rr][0] = need[rr] ? sx + dhz[rr][0] : 0.f;
        carry[rr][1] = need[rr] ? sy + dhz[rr][1] : 0.f;
      }
    }
    half_barrier_sync(1 + mt);                                          // last step's reads of the A tile are done
    // ---- gate gradients of the own (row, unit) pairs
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int b = b0 + row0 + rr * 8;
      float drp[2] = {0.f, 0.f}, dzp[2] = {0.f, 0.f}, dnp[2] = {0.f, 0.f}, dnr[2] = {0.f, 0.f};
      if (b < B) {
        const long long bt = (long long)b * T + tt;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float r = u ? r2[rr].y : r2[rr].x, z = u ? z2[rr].y : z2[rr].x, n = u ? n2[rr].y : n2[rr].x;
          const float ghn = u ? g2[rr].y : g2[rr].x, hp = u ? hp2[rr].y : hp2[rr].x;
          const float dh = (u ? do2[rr].y : do2[rr].x) + carry[rr][u];
          dnp[u] = dh * (1.f - z) * (1.f - n * n);
          dzp[u] = dh * (hp - n) * z * (1.f - z);
          drp[u] = dnp[u] * ghn * r * (1.f - r);
          dnr[u] = dnp[u] * r;
          dhz[rr][u] = dh * z;
          bsum[0][u] += drp[u]; bsum[1][u] += dzp[u]; bsum[2][u] += dnp[u]; bsum[3][u] += dnr[u];
        }
        float* gx = dGx + (bt * 2 + d) * 3 * H + j;
        float* gh = dGh + (bt * 2 + d) * 3 * H + j;
        *reinterpret_cast<float2*>(gx) = make_float2(drp[0], drp[1]);
        *reinterpret_cast<float2*>(gx + H) = make_float2(dzp[0], dzp[1]);
        *reinterpret_cast<float2*>(gx + 2 * H) = make_float2(dnp[0], dnp[1]);
        *reinterpret_cast<float2*>(gh) = make_float2(drp[0], drp[1]);
        *reinterpret_cast<float2*>(gh + H) = make_float2(dzp[0], dzp[1]);
        *reinterpret_cast<float2*>(gh + 2 * H) = make_float2(dnr[0], dnr[1]);
        *reinterpret_cast<float2*>(Hprev + (bt * 2 + d) * H + j) = hp2[rr];
        if (dGx16) {      // bf16 copies for the weight-gradient GEMMs that follow (dW_ih = dGx^T X, dW_hh = dGh^T Hprev)
          __nv_bfloat16* gx16 = dGx16 + (bt * 2 + d) * 3 * H + j;
          __nv_bfloat16* gh16 = dGh16 + (bt * 2 + d) * 3 * H + j;
          const uint32_t pr = pack_bf16x2(drp[0], drp[1]), pz = pack_bf16x2(dzp[0], dzp[1]);
          *reinterpret_cast<uint32_t*>(gx16) = pr;
          *reinterpret_cast<uint32_t*>(gx16 + H) = pz;
          *reinterpret_cast<uint32_t*>(gx16 + 2 * H) = pack_bf16x2(dnp[0], dnp[1]);
          *reinterpret_cast<uint32_t*>(gh16) = pr;
          *reinterpret_cast<uint32_t*>(gh16 + H) = pz;
          *reinterpret_cast<uint32_t*>(gh16 + 2 * H) = pack_bf16x2(dnr[0], dnr[1]);
          *reinterpret_cast<uint32_t*>(Hprev16 + (bt * 2 + d) * H + j) = pack_bf16x2(hp2[rr].x, hp2[rr].y);
        }
      }
      // the A operand of this step's product: dGh of (row, gate, own unit), zero for rows beyond the batch
      const int ar = (row0 + rr * 8) * kLdB + jl;
      uint32_t hi, lo;
      split2(drp[0], drp[1], hi, lo);
      *reinterpret_cast<uint32_t*>(sAh + ar) = hi;
      *reinterpret_cast<uint32_t*>(sAl + ar) = lo;
      split2(dzp[0], dzp[1], hi, lo);
      *reinterpret_cast<uint32_t*>(sAh + ar + kPJ) = hi;
      *reinterpret_cast<uint32_t*>(sAl + ar + kPJ) = lo;
      split2(dnr[0], dnr[1], hi, lo);
      *reinterpret_cast<uint32_t*>(sAh + ar + 2 * kPJ) = hi;
      *reinterpret_cast<uint32_t*>(sAl + ar + 2 * kPJ) = lo;
    }
    if (bs == T - 1) break;                                             // dh of the step before the first is not needed
    half_barrier_sync(1 + mt);                                          // the half's A tile is complete
    // ---- partial[16 rows][64 columns oct * 64 ..] = A (16 x 96) . W_hh[own rows][columns]
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const __nv_bfloat16* ah = sAh + row0 * kLdB + tig * 2;
    const __nv_bfloat16* al = sAl + row0 * kLdB + tig * 2;
    const __nv_bfloat16* wh = sWh + (oct * 64 + gid) * kLdB + tig * 2;
    const __nv_bfloat16* wl = sWl + (oct * 64 + gid) * kLdB + tig * 2;
#pragma unroll
    for (int kk = 0; kk < 3 * kPJ; kk += 16) {
      uint32_t a_hi[4], a_lo[4];
      a_hi[0] = *reinterpret_cast<const uint32_t*>(ah + kk);
      a_hi[1] = *reinterpret_cast<const uint32_t*>(ah + 8 * kLdB + kk);
      a_hi[2] = *reinterpret_cast<const uint32_t*>(ah + kk + 8);
      a_hi[3] = *reinterpret_cast<const uint32_t*>(ah + 8 * kLdB + kk + 8);
      a_lo[0] = *reinterpret_cast<const uint32_t*>(al + kk);
      a_lo[1] = *reinterpret_cast<const uint32_t*>(al + 8 * kLdB + kk);
      a_lo[2] = *reinterpret_cast<const uint32_t*>(al + kk + 8);
      a_lo[3] = *reinterpret_cast<const uint32_t*>(al + 8 * kLdB + kk + 8);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t b_hi[2], b_lo[2];
        b_hi[0] = *reinterpret_cast<const uint32_t*>(wh + nt * 8 * kLdB + kk);
        b_hi[1] = *reinterpret_cast<const uint32_t*>(wh + nt * 8 * kLdB + kk + 8);
        b_lo[0] = *reinterpret_cast<const uint32_t*>(wl + nt * 8 * kLdB + kk);
        b_lo[1] = *reinterpret_cast<const uint32_t*>(wl + nt * 8 * kLdB + kk + 8);
        mma_bf16_m16n8k16(acc[nt], a_hi, b_hi);
        mma_bf16_m16n8k16(acc[nt], a_lo, b_hi);
        mma_bf16_m16n8k16(acc[nt], a_hi, b_lo);
      }
    }
    // ---- the partial sums travel to the CTAs that own their columns (accumulator element i: row gid + 8 (i >> 1),
    //      column oct * 64 + nt * 8 + tig * 2 + (i & 1))
    {
      const uint32_t tg = (uint32_t)(bs + 1);
      uint4* xs = xbase + ((bs & 1) * 2 + mt) * kPGroup * kPGroup * kPairVecs + r_cta * kPairVecs + gid * 16 + tig;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int col = oct * 64 + nt * 8;                              // first column of the n tile
        uint4* xc = xs + (col >> 5) * kPGroup * kPairVecs + ((col & 31) >> 1);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
          if (b0 + row0 + rr * 8 < B)
            st_relaxed_v4(xc + rr * 8 * 16, __float_as_uint(acc[nt][rr * 2]), tg, __float_as_uint(acc[nt][rr * 2 + 1]), tg);
      }
    }
  }
  // ---- bias gradients db_ih = sum dGx, db_hh = sum dGh over (batch, time): this thread's two rows are summed, the
  //      eight row lanes of the warp follow (fixed shuffle tree), one partial row per (batch tile, half):
  //      bias_partial[(tile * 2 + half)][ih | hh][direction][3H]   (summed in row order by sed_reduce_partials)
  if (bias_partial != nullptr) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float v = bsum[k][u];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        bsum[k][u] = v;
      }
    if (gid == 0) {
      const long long prow = (long long)(q.b_base / kPB + blockIdx.y) * 2 + mt;
      float* pih = bias_partial + ((prow * 2 + 0) * 2 + d) * 3 * H + j;
      float* phh = bias_partial + ((prow * 2 + 1) * 2 + d) * 3 * H + j;
      *reinterpret_cast<float2*>(pih) = make_float2(bsum[0][0], bsum[0][1]);
      *reinterpret_cast<float2*>(pih + H) = make_float2(bsum[1][0], bsum[1][1]);
      *reinterpret_cast<float2*>(pih + 2 * H) = make_float2(bsum[2][0], bsum[2][1]);
      *reinterpret_cast<float2*>(phh) = make_float2(bsum[0][0], bsum[0][1]);
      *reinterpret_cast<float2*>(phh + H) = make_float2(bsum[1][0], bsum[1][1]);
      *reinterpret_cast<float2*>(phh + 2 * H) = make_float2(bsum[3][0], bsum[3][1]);
    }
  }
}

// Cooperative launches (co-residency of a group is what makes the barrier deadlock-free) over chunks of
// batch tiles that fit the device at one CTA per SM.
// exchange area of one batch tile: 2 directions x 2 slots x 32 rows x (values per row) 8-byte words
constexpr size_t gru_xchg_tile_bytes(int values_per_row) { return (size_t)2 * 2 * kPB * values_per_row * 8; }

template <typename F>
int launch_gru_persistent(const void* kern, const char* name, size_t smem, int B, int T, void* xchg_ws,
                          int values_per_row, cudaStream_t stream, F fill_args) {
  SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SED_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPThreads, smem));
  const int max_tiles = per_sm * sm_count() / (kPGroup * 2);
  SED_REQUIRE(max_tiles >= 1, "%s: the device cannot hold one batch tile", name);
  const int tiles = (B + kPB - 1) / kPB;
  const size_t tile_bytes = gru_xchg_tile_bytes(values_per_row);
  SED_REQUIRE((reinterpret_cast<uintptr_t>(xchg_ws) & 15) == 0, "%s: the workspace must be 16-byte aligned", name);
  SED_CUDA(cudaMemsetAsync(xchg_ws, 0, tile_bytes * (size_t)tiles, stream));      // tag 0 = nothing produced yet
  for (int t0 = 0; t0 < tiles; t0 += max_tiles) {
    const int nt = tiles - t0 < max_tiles ? tiles - t0 : max_tiles;
    GruPersist q{reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(xchg_ws) + tile_bytes * (size_t)t0), B, T, t0 * kPB};
    void* args[12];
    const int n = fill_args(args);
    args[n] = &q;
    SED_CUDA(cudaLaunchCooperativeKernel(kern, dim3(kPGroup, nt, 2), dim3(kPThreads), args, smem, stream));
    count_launch();
  }
  return 0;
}

// (R, K) fp32 -> (R, 3K) bf16:  which = 0: [hi | lo | hi]   (left operand)
//                               which = 1: [hi | hi | lo]   (right operand)
// so that  A' . B'^T = hi.hi + lo.hi + hi.lo  ~  fp32-accurate product on the bf16 tensor cores.
__global__ void split_bf16x3_kernel(const float* __restrict__ x, long long R, int K, int which,
                                    __nv_bfloat16* __restrict__ out) {
  const long long n = R * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / K;
    const int k = (int)(i % K);
    const float v = x[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * 3 * K;
    o[k] = hi;
    o[K + k] = which == 0 ? lo : hi;
    o[2 * K + k] = which == 0 ? hi : lo;
  }
}

// (R, C) fp32 -> (C, R) bf16 (tiled through shared memory)
__global__ void transpose_to_bf16_kernel(const float* __restrict__ x, int R, int C, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? x[(long long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(long long)c * R + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// column sums of a (rows, C) fp32 matrix: partial[blk][C]
__global__ void colsum_kernel(const float* __restrict__ x, long long rows, int C, float* __restrict__ partial) {
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (long long r = r0; r < r1; ++r) a += x[r * C + c];
    partial[(long long)blockIdx.x * C + c] = a;
  }
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

long long sed_gru_workspace_bytes(int B, int H, int backward) {
  const long long tiles = (B + kPB - 1) / kPB;
  const long long step_kernels = backward ? 4LL * B * H * (long long)sizeof(float) : 16;   // (2,2,B,H) carry / nothing
  if (H != kPH) return step_kernels;
  const long long persistent = tiles * (long long)gru_xchg_tile_bytes(backward ? kBwdXchgValues : kPH);
  return persistent > step_kernels ? persistent : step_kernels;
}

int sed_gru_fwd(const float* gx, const float* w_hh, const float* b_hh, float* out, float* gates, void* sync_ws,
                int B, int T, int H, sed_stream_t stream) {
  SED_REQUIRE(gx && w_hh && b_hh && out && gates, "sed_gru_fwd: null pointer");
  SED_REQUIRE(H % kHT == 0 && H % 4 == 0 && H <= 512, "sed_gru_fwd: hidden size %d unsupported", H);
  if (B == 0 || T == 0) return 0;
  if (H == kPH && sync_ws != nullptr) {
    const size_t psmem = sizeof(__nv_bfloat16) * (size_t)(2 * 3 * kPJ * kLdK + 2 * kPB * kLdK);
    return launch_gru_persistent((const void*)gru_fwd_persistent_kernel, "sed_gru_fwd", psmem, B, T, sync_ws, kPH,
                                 (cudaStream_t)stream, [&](void** a) {
                                   a[0] = (void*)&gx; a[1] = (void*)&w_hh; a[2] = (void*)&b_hh; a[3] = (void*)&out; a[4] = (void*)&gates;
                                   return 5;
                                 });
  }
  const size_t smem = sizeof(float) * (size_t)(3 * kHT * (H + 4) + kBT * H);
  SED_CUDA(cudaFuncSetAttribute(gru_fwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid(H / kHT, (B + kBT - 1) / kBT, 2);
  for (int s = 0; s < T; ++s) {
    gru_fwd_step_kernel<<<grid, kGruThreads, smem, (cudaStream_t)stream>>>(gx, w_hh, b_hh, out, gates, B, T, H, s);
    SED_LAUNCH_CHECK("gru_fwd_step_kernel");
  }
  return 0;
}

int sed_gru_bwd_bias_rows(int B) { return 2 * ((B + kPB - 1) / kPB); }

int sed_gru_bwd(const float* dout, const float* out, const float* gates, const float* w_hh, float* carry /* (2,2,B,H) */,
                float* dgx, float* dgh, float* hprev, void* dgx_bf16, void* dgh_bf16, void* hprev_bf16, float* bias_partial,
                int B, int T, int H, sed_stream_t stream) {
  SED_REQUIRE(dout && out && gates && w_hh && carry && dgx && dgh && hprev, "sed_gru_bwd: null pointer");
  const bool want16 = dgx_bf16 || dgh_bf16 || hprev_bf16;
  SED_REQUIRE(!want16 || (dgx_bf16 && dgh_bf16 && hprev_bf16 && H == kPH),
              "sed_gru_bwd: the bf16 copies come all three or not at all, and only from the H = %d kernel", kPH);
  SED_REQUIRE(H % kHT == 0 && H % 4 == 0 && H <= 256, "sed_gru_bwd: hidden size %d unsupported", H);
  SED_REQUIRE(bias_partial == nullptr || H == kPH, "sed_gru_bwd: the bias-gradient partials come only from the H = %d kernel",
              kPH);
  if (B == 0 || T == 0) return 0;
  if (H == kPH) {
    // dh never leaves registers; the scratch is the exchange area of the partial sums
    const size_t psmem = sizeof(__nv_bfloat16) * (size_t)(2 * kPH * kLdB + 2 * kPB * kLdB);
    return launch_gru_persistent((const void*)gru_bwd_persistent_kernel, "sed_gru_bwd", psmem, B, T, carry, kBwdXchgValues,
                                 (cudaStream_t)stream, [&](void** a) {
                                   a[0] = (void*)&dout; a[1] = (void*)&out; a[2] = (void*)&gates; a[3] = (void*)&w_hh;
                                   a[4] = (void*)&dgx; a[5] = (void*)&dgh; a[6] = (void*)&hprev;
                                   a[7] = (void*)&dgx_bf16; a[8] = (void*)&dgh_bf16; a[9] = (void*)&hprev_bf16;
                                   a[10] = (void*)&bias_partial;
                                   return 11;
                                 });
  }
  const size_t smem = sizeof(float) * (size_t)(3 * H * kHT + kBT * 3 * H + kBT * kHT);
  SED_CUDA(cudaFuncSetAttribute(gru_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid(H / kHT, (B + kBT - 1) / kBT, 2);
  const long long half = 2LL * B * H;
  for (int bs = 0; bs < T; ++bs) {
    const float* cin = carry + (bs & 1) * half;
    float* cout = carry + ((bs + 1) & 1) * half;
    gru_bwd_step_kernel<<<grid, kGruThreads, smem, (cudaStream_t)stream>>>(dout, out, gates, w_hh, cin, cout, dgx, dgh,
                                                                           hprev, B, T, H, bs);
    SED_LAUNCH_CHECK("gru_bwd_step_kernel");
  }
  return 0;
}

int sed_split_bf16x3(const float* x, long long R, int K, int which, void* out, sed_stream_t stream) {
  SED_REQUIRE(x && out && (which == 0 || which == 1), "sed_split_bf16x3: bad arguments");
  if (R * K == 0) return 0;
  const int grid = (int)min((R * K + 255) / 256, (long long)sm_count() * 8);
  split_bf16x3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, R, K, which, reinterpret_cast<__nv_bfloat16*>(out));
  SED_LAUNCH_CHECK("split_bf16x3_kernel");
  return 0;
}

int sed_transpose_to_bf16(const float* x, int R, int C, void* out, sed_stream_t stream) {
  SED_REQUIRE(x && out && R > 0 && C > 0, "sed_transpose_to_bf16: bad arguments");
  const dim3 grid((C + 31) / 32, (R + 31) / 32);
  transpose_to_bf16_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, R, C,
                                                                           reinterpret_cast<__nv_bfloat16*>(out));
  SED_LAUNCH_CHECK("transpose_to_bf16_kernel");
  return 0;
}

int sed_colsum_f32(const float* x, long long rows, int C, float* partial, sed_stream_t stream) {
  SED_REQUIRE(x && partial && rows >= 1 && C >= 1, "sed_colsum_f32: bad arguments");
  colsum_kernel<<<sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(x, rows, C, partial);
  SED_LAUNCH_CHECK("colsum_kernel");
  return 0;
}

}  // extern "C"
