// conv_halo2_tc.cu -- the halo-tile 3x3 convolution (see conv_halo_tc.cu) on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// Forward and data gradient of ConvBlock.conv1 / conv2 (/root/reference/pytorch/models.py:102-103, backward at
// pytorch/main.py:257) for every layer with Cin % 64 == 0.  The single-CTA kernel is bound by shared-memory bandwidth:
// per 128-cycle M128 x N256 x K16 MMA an SM reads 4 KB of A and 8 KB of B out of shared memory and TMA writes the
// same 12 KB per K step into it.  A CTA pair issues ONE M = 256 MMA per step instead: each CTA keeps its own 128
// pixels of A and only HALF of the weight tile (128 of the 256 output channels), the tensor cores of both SMs read
// both halves.  Per SM that is 8 KB instead of 12 KB of operand reads and 16 KB instead of 32 KB of weight bytes
// through TMA per tap.
//
// Pair tile = 2 x kMT vertically adjacent 128-pixel M tiles (cluster rank r owns kMT of them: image rows
// h0 + r*kMT*bh ...) x kN output channels; kN = 256 / kMT = 1, kN = 128 or 64 / kMT = 2 (TMEM: 2 x kMT x kN columns).
// For 64 -> 64 channels each CTA keeps its half of the whole 9-tap weight set resident.  Roles per CTA as in the single-CTA kernel (warp 0 TMA producer, warp 1 MMA issuer -- only the leader
// CTA, cluster rank 0, issues -- warps 2-9 epilogue).  Barriers:
//   a_full / b_full   leader's copies: one arrive.expect_tx by the leader's producer for the bytes of BOTH CTAs, the
//                     peer's TMA completes its bytes on the leader's barrier;
//   a_empty / b_empty / tmem_full   one copy per CTA, released by the leader's tcgen05.commit multicast to both CTAs;
//   tmem_empty        leader's copy, 2 x 256 epilogue threads arrive (the peer's remotely).
#include "common.cuh"
#include "tc.cuh"

namespace sed {
namespace {

using namespace tc;

constexpr int k2Threads = 320;
constexpr int k2Epi = 256;
constexpr int k2MaxA = 4, k2MaxB = 8;

struct Halo2Params {
  int B, H, W, Cin, Cout;
  int bh;              // image rows per M tile (128 / W); a pair tile covers 2 * kMT * bh rows
  int rows_cta;        // kMT * bh: image rows owned by one CTA of the pair
  int tiles_h, tiles_n, num_tiles;
  int kb_per_tap;
  int a_bytes;         // (rows_cta + 2) * W * 128
  int a_stages, b_stages;
  __nv_bfloat16* y;
  float* stats;        // [gridDim.x * 4][2][Cout] or nullptr
};

__device__ __forceinline__ void warp_transpose_reduce2b(float (&a)[32], float (&b)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? a[i] : a[i + off];
      const float keep = hi ? a[i + off] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      const float send2 = hi ? b[i] : b[i + off];
      const float keep2 = hi ? b[i + off] : b[i];
      b[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
    }
  }
}

template <int k2N, int kMT, bool kResidentB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
conv3x3_halo2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const Halo2Params p) {
  constexpr int k2BBytes = (k2N / 2) * 128;          // this CTA's half of one weight tap tile
  constexpr uint32_t k2TmemCols = 2 * kMT * k2N;
  static_assert(k2TmemCols == 256 || k2TmemCols == 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.a_stages * p.a_bytes;
  __shared__ uint64_t a_full[k2MaxA], a_empty[k2MaxA], b_full[k2MaxB], b_empty[k2MaxB];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2], b_res_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < k2MaxA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < k2MaxB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 2 * k2Epi); }
    mbar_init(&b_res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<k2TmemCols>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs' barriers are initialised before any remote signal
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      if (kResidentB) {                     // each CTA: its half of all 9 * kb_per_tap weight tiles, once
        const int nb = 9 * p.kb_per_tap;
        if (leader) mbar_arrive_expect_tx(&b_res_full, 2u * (uint32_t)(nb * k2BBytes));
        for (int i = 0; i < nb; ++i)
          tma_load_2d_2cta(smem_b + i * k2BBytes, &tmap_b, &b_res_full, i * 64, (int)rank * (k2N / 2));
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
        const int n_tile = tile % p.tiles_n;
        const int m_tile = tile / p.tiles_n;
        const int b = m_tile / p.tiles_h;
        const int h0 = (m_tile % p.tiles_h) * 2 * p.rows_cta + (int)rank * p.rows_cta;
        for (int kw = 0; kw < 3; ++kw) {
          for (int cb = 0; cb < p.kb_per_tap; ++cb) {
            mbar_wait(&a_empty[as], aph ^ 1);
            if (leader) mbar_arrive_expect_tx(&a_full[as], 2u * (uint32_t)p.a_bytes);
            tma_load_4d_2cta(smem_a + as * p.a_bytes, &tmap_a, &a_full[as], cb * 64, kw - 1, h0 - 1, b);
            if (++as == p.a_stages) { as = 0; aph ^= 1; }
            if (kResidentB) continue;
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (leader) mbar_arrive_expect_tx(&b_full[bs], 2u * (uint32_t)k2BBytes);
              tma_load_2d_2cta(smem_b + bs * k2BBytes, &tmap_b, &b_full[bs], (kh * 3 + kw) * p.Cin + cb * 64,
                               n_tile * k2N + (int)rank * (k2N / 2));
              if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, k2N, 0, 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      if (kResidentB) {
        mbar_wait(&b_res_full, 0);
        tcgen05_fence_after();
      }
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (kMT * k2N);
        for (int kw = 0; kw < 3; ++kw) {
          for (int cb = 0; cb < p.kb_per_tap; ++cb) {
            mbar_wait(&a_full[as], aph);
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem_a + as * p.a_bytes);
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
              uint32_t sb;
              if (kResidentB) {
                sb = smem_u32(smem_b + ((kh * 3 + kw) * p.kb_per_tap + cb) * k2BBytes);
              } else {
                mbar_wait(&b_full[bs], bph);
                tcgen05_fence_after();
                sb = smem_u32(smem_b + bs * k2BBytes);
              }
              const uint32_t first = (kw | cb | kh) == 0 ? 0u : 1u;
#pragma unroll
              for (int mt = 0; mt < kMT; ++mt) {
                const uint32_t sam = sa + (uint32_t)((kh + mt * p.bh) * p.W) * 128u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t da = umma_desc_sw128(sam + ks * 32, 16, 1024);
                  const uint64_t db = umma_desc_sw128(sb + ks * 32, 16, 1024);
                  umma_bf16_2cta(d_tmem + mt * k2N, da, db, idesc, (first | (uint32_t)ks) != 0 ? 1u : 0u);
                }
              }
              if (!kResidentB) {
                umma_commit_2cta(&b_empty[bs]);
                if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
              }
            }
            umma_commit_2cta(&a_empty[as]);
            if (++as == p.a_stages) { as = 0; aph ^= 1; }
          }
        }
        umma_commit_2cta(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int r_h = row / p.W, r_w = row % p.W;
    constexpr bool kLocalStats = k2N == 64;
    constexpr int kChunksPerWarp = k2N / 64;
    float* stat_row = p.stats == nullptr ? nullptr : p.stats + ((long long)blockIdx.x * 4 + q) * 2 * p.Cout;
    if (stat_row != nullptr && !kLocalStats) {
      for (int nt = 0; nt < p.tiles_n; ++nt)
#pragma unroll
        for (int j = 0; j < kChunksPerWarp; ++j) {
          const int ch = nt * k2N + (half + 2 * j) * 32 + lane;
          stat_row[ch] = 0.f;
          stat_row[p.Cout + ch] = 0.f;
        }
    }
    float s_acc[32], ss_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s_acc[i] = ss_acc[i] = 0.f;
    int it = 0;
    for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++it) {
      const int acc = it & 1;
      const int n_tile = tile % p.tiles_n;
      const int m_tile = tile / p.tiles_n;
      const int b = m_tile / p.tiles_h;
      const int h0 = (m_tile % p.tiles_h) * 2 * p.rows_cta + (int)rank * p.rows_cta;
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int j = 0; j < kChunksPerWarp; ++j) {
        const int c = half + 2 * j;
#pragma unroll 1
        for (int mt = 0; mt < kMT; ++mt) {
          const int h = h0 + mt * p.bh + r_h;
          const bool valid = h < p.H;
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (acc * kMT + mt) * k2N + c * 32, v);
          if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(p.y + (((long long)b * p.H + h) * p.W + r_w) * p.Cout + n_tile * k2N + c * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
              o.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
              o.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
              o.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
              d4[g] = o;
            }
            if (stat_row != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                s_acc[i] += v[i];
                ss_acc[i] = fmaf(v[i], v[i], ss_acc[i]);
              }
            }
          }
        }
        if (!kLocalStats && stat_row != nullptr) {
          warp_transpose_reduce2b(s_acc, ss_acc, lane);
          const int ch = n_tile * k2N + c * 32 + lane;
          stat_row[ch] += s_acc[0];
          stat_row[p.Cout + ch] += ss_acc[0];
#pragma unroll
          for (int i = 0; i < 32; ++i) s_acc[i] = ss_acc[i] = 0.f;
        }
      }
      tcgen05_fence_before();
      mbar_arrive_leader(&tmem_empty_bar[acc]);
    }
    if (kLocalStats && stat_row != nullptr) {          // k2N == 64: Cout == 64, a single n tile; one reduce per warp
      warp_transpose_reduce2b(s_acc, ss_acc, lane);
      stat_row[half * 32 + lane] = s_acc[0];
      stat_row[p.Cout + half * 32 + lane] = ss_acc[0];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                       // no CTA of the pair exits while its peer may still signal its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta<k2TmemCols>(tmem_base);
  }
}

}  // namespace
}  // namespace sed

namespace sed {
namespace {

struct Pair2Plan { int kN, kMT, resident, a_stages, b_stages, smem, grid; };

bool make_plan2(const Halo2Params& p, Pair2Plan* q) {
  q->kN = p.Cout == 64 ? 64 : (p.Cout == 128 ? 128 : 256);
  q->kMT = q->kN == 256 ? 1 : 2;
  q->resident = (p.Cout == 64 && p.Cin == 64) ? 1 : 0;
  const int b_bytes = (q->kN / 2) * 128;
  const int budget = 220 * 1024;
  if (q->resident) {
    const int res = 9 * p.kb_per_tap * b_bytes;
    q->a_stages = (budget - res) / p.a_bytes;
    if (q->a_stages > k2MaxA) q->a_stages = k2MaxA;
    q->b_stages = 0;
    q->smem = res + q->a_stages * p.a_bytes + 1024;
  } else {
    q->a_stages = 3;
    q->b_stages = (budget - q->a_stages * p.a_bytes) / b_bytes;
    if (q->b_stages > k2MaxB) q->b_stages = k2MaxB;
    q->smem = q->a_stages * p.a_bytes + q->b_stages * b_bytes + 1024;
  }
  return q->a_stages >= 2 && (q->resident || q->b_stages >= 3);
}

template <int kN, int kMT, bool kRes>
int launch_halo2(const CUtensorMap& ta, const CUtensorMap& tb, const Halo2Params& p, const Pair2Plan& q,
                 cudaStream_t stream) {
  auto kern = conv3x3_halo2_kernel<kN, kMT, kRes>;
  SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, q.smem));
  kern<<<q.grid, k2Threads, q.smem, stream>>>(ta, tb, p);
  SED_LAUNCH_CHECK("conv3x3_halo2_kernel");
  return 0;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_conv3x3_tc2_grid(int B, int H, int W, int Cin, int Cout) {   // rows of the statistics workspace (4 per CTA)
  (void)B; (void)H; (void)W; (void)Cin; (void)Cout;
  return (sm_count() & ~1) * 4;
}

int sed_conv3x3_tc2_fwd(const void* x, const void* wpack, void* y, float* stats_partial, int B, int H, int W, int Cin,
                        int Cout, sed_stream_t stream) {
  SED_REQUIRE(x && wpack && y, "sed_conv3x3_tc2_fwd: null pointer");
  SED_REQUIRE(B > 0 && H > 0, "sed_conv3x3_tc2_fwd: empty batch");
  SED_REQUIRE(W >= 8 && W <= 128 && 128 % W == 0, "sed_conv3x3_tc2_fwd: W=%d must divide 128 and be >= 8", W);
  SED_REQUIRE(Cin % 64 == 0 && Cin >= 64, "sed_conv3x3_tc2_fwd: Cin=%d must be a multiple of 64", Cin);
  SED_REQUIRE((Cout == 64 || Cout == 128 || Cout % 256 == 0) && Cout <= 512, "sed_conv3x3_tc2_fwd: Cout=%d unsupported", Cout);
  SED_REQUIRE(aligned(y, 16), "sed_conv3x3_tc2_fwd: output must be 16-byte aligned");
  Halo2Params p;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bh = 128 / W;
  p.kb_per_tap = Cin / 64;
  Pair2Plan q;
  {
    const int kN = Cout == 64 ? 64 : (Cout == 128 ? 128 : 256);
    p.rows_cta = (kN == 256 ? 1 : 2) * p.bh;
    p.tiles_n = Cout / kN;
  }
  p.a_bytes = (p.rows_cta + 2) * W * 128;
  SED_REQUIRE(make_plan2(p, &q), "sed_conv3x3_tc2_fwd: no tiling for W=%d Cin=%d Cout=%d", W, Cin, Cout);
  p.tiles_h = (H + 2 * p.rows_cta - 1) / (2 * p.rows_cta);
  const long long tiles = (long long)B * p.tiles_h * p.tiles_n;
  SED_REQUIRE(tiles < (1LL << 31), "sed_conv3x3_tc2_fwd: too many tiles");
  p.num_tiles = (int)tiles;
  p.a_stages = q.a_stages; p.b_stages = q.b_stages;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.stats = stats_partial;

  alignas(64) CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)(p.rows_cta + 2), 1};
    if (int rc = tc::make_tmap_bf16(&ta, x, 4, dims, strides, box, "conv2 activation halo map")) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout};
    const uint64_t strides[1] = {(uint64_t)9 * Cin * 2};
    const uint32_t box[2] = {64, (uint32_t)(q.kN / 2)};
    if (int rc = tc::make_tmap_bf16(&tb, wpack, 2, dims, strides, box, "conv2 weight map")) return rc;
  }
  q.grid = sm_count() & ~1;
  if ((long long)q.grid > 2 * tiles) q.grid = (int)(2 * tiles);
  // every statistics row of the workspace must be defined: rows of CTAs beyond a shortened grid are zeroed here
  if (stats_partial != nullptr && q.grid < (sm_count() & ~1))
    SED_CUDA(cudaMemsetAsync(stats_partial, 0, sizeof(float) * (size_t)sed_conv3x3_tc2_grid(B, H, W, Cin, Cout) * 2 * Cout,
                             (cudaStream_t)stream));
  cudaStream_t s = (cudaStream_t)stream;
  if (q.kN == 64 && q.resident) return launch_halo2<64, 2, true>(ta, tb, p, q, s);
  if (q.kN == 64) return launch_halo2<64, 2, false>(ta, tb, p, q, s);
  if (q.kN == 128) return launch_halo2<128, 2, false>(ta, tb, p, q, s);
  return launch_halo2<256, 1, false>(ta, tb, p, q, s);
}

}  // extern "C"
