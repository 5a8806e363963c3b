// conv_wgrad_tc.cu -- weight gradient of the 3x3 convolution as a tcgen05 split-K GEMM (sm_100a).
//
// Replaces cuDNN's wgrad behind loss.backward() for ConvBlock.conv1/conv2
// (/root/reference/pytorch/main.py:257 through /root/reference/pytorch/models.py:102-103).
//
//   dW[tap][co][ci] = sum over pixels p of  dY[p][co] * X[p + tap_offset][ci]
//
// Per tap this is a GEMM with M = Cout, N = Cin and K = B*H*W pixels.  Both operands are NHWC
// bf16, i.e. the *pixel* (K) index is the strided one and the channel (M / N) index is
// contiguous: "MN-major" operands.  A TMA box {64 ch, W, bh, 1} lands in shared memory as
// 128 pixel rows x 128 B with the 128B swizzle -- exactly the canonical MN-major SW128 slab that
// tcgen05.mma reads with (LBO = slab stride, SBO = 1024 B), so no transpose is ever done.
// The tap shift (kh-1, kw-1) is applied to the X box coordinates; TMA zero-fills the halo.
//
// grid = 9 taps x (Cout/128) x (Cin/kN) x splits; each CTA accumulates its K range in TMEM and
// writes one fp32 slab [tap][co][ci]; sed_conv_unpack_wgrad sums the split slabs in a fixed
// order (deterministic, no atomics) while converting to OIHW.
#include "common.cuh"
#include "tc.cuh"

namespace sed {
namespace {

using namespace tc;

constexpr int kSlabBytes = 128 * 128;     // 128 pixel rows x 64 bf16
constexpr int kThreadsW = 192;

template <int kN> struct WCfg {
  static constexpr int kBSlabs = kN / 64;
  static constexpr int kStageBytes = (2 + kBSlabs) * kSlabBytes;
  static constexpr int kStages = (kN == 128) ? 3 : 4;
  static constexpr uint32_t kTmemCols = kN;           // 64 or 128 (power of two >= 32)
  static constexpr int kDynSmem = kStages * kStageBytes + 1024;
};

struct WgradParams {
  int B, H, W, Cin, Cout;
  int bh, tiles_h;
  int m_tiles, n_tiles, splits;
  int a_slabs;             // 2, or 1 when Cout == 64 (upper 64 accumulator rows are ignored)
  int gemm;                // 1: plain out[M][N] = A[R][M]^T * B[R][N] (one "tap", rows tiled by 128)
  int k_total;             // B * tiles_h pixel tiles
  float* out;              // [splits][9][Cout][Cin]
  long long slab_stride;   // 9*Cout*Cin
};

template <int kN>
__global__ void __launch_bounds__(kThreadsW, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                        const WgradParams p) {
  using C = WCfg<kN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[C::kStages], empty_bar[C::kStages], done_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int rest = blockIdx.x;
  const int split = rest % p.splits; rest /= p.splits;
  const int n_tile = rest % p.n_tiles; rest /= p.n_tiles;
  const int m_tile = rest % p.m_tiles;
  const int tap = rest / p.m_tiles;
  const int kh = tap / 3, kw = tap % 3;
  const int per = (p.k_total + p.splits - 1) / p.splits;
  const int k_begin = split * per;
  const int k_end = min(p.k_total, k_begin + per);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)(p.a_slabs + C::kBSlabs) * kSlabBytes;
      for (int kt = k_begin; kt < k_end; ++kt) {
        const int b = kt / p.tiles_h;
        const int h0 = (kt % p.tiles_h) * p.bh;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::kStageBytes;
        uint8_t* sb = sa + 2 * kSlabBytes;
        mbar_arrive_expect_tx(&full_bar[stage], bytes);
        if (p.gemm) {
          for (int s = 0; s < p.a_slabs; ++s)
            tma_load_4d(sa + s * kSlabBytes, &tmap_dy, &full_bar[stage], m_tile * 128 + s * 64, kt * 128, 0, 0);
#pragma unroll
          for (int s = 0; s < C::kBSlabs; ++s)
            tma_load_4d(sb + s * kSlabBytes, &tmap_x, &full_bar[stage], n_tile * kN + s * 64, kt * 128, 0, 0);
        } else {
          for (int s = 0; s < p.a_slabs; ++s)
            tma_load_4d(sa + s * kSlabBytes, &tmap_dy, &full_bar[stage], m_tile * 128 + s * 64, 0, h0, b);
#pragma unroll
          for (int s = 0; s < C::kBSlabs; ++s)
            tma_load_4d(sb + s * kSlabBytes, &tmap_x, &full_bar[stage], n_tile * kN + s * 64, kw - 1, h0 + kh - 1, b);
        }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kN, 1, 1);
      const uint32_t a_lbo = p.a_slabs == 2 ? kSlabBytes : 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k_begin; kt < k_end; ++kt) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
        const uint32_t sb = sa + 2 * kSlabBytes;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {           // 128 pixels = 8 x K16
          const uint64_t da = umma_desc_sw128(sa + ks * 2048, a_lbo, 1024);
          const uint64_t db = umma_desc_sw128(sb + ks * 2048, kSlabBytes, 1024);
          umma_bf16(tmem_base, da, db, idesc, (kt > k_begin || ks > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    const int q = warp & 3;
    const int co_local = q * 32 + lane;
    const int co = m_tile * 128 + co_local;
    float* dst = p.out + (long long)split * p.slab_stride + ((long long)tap * p.Cout + co) * p.Cin + n_tile * kN;
    const bool valid = co < p.Cout && co_local < p.a_slabs * 64;
    if (k_end > k_begin) {
      mbar_wait(&done_bar, 0);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < kN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, v);
        if (valid) {
          float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) d4[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
      }
    } else if (valid) {
      for (int c = 0; c < kN; ++c) dst[c] = 0.f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient, "kw-group" tiling (the kernel the 3x3 layers use).
//
//   dW[kh][kw][co][ci] = sum_{b,h,w'} dY[b][h][w' - (kw-1)][co] * X[b][h + kh - 1][w'][ci]
//
// A CTA owns one kw (or, for Cout = 64, a pair of kw stacked in the M dimension), a 128-row co tile,
// a kN-column ci tile and a split of the pixel range, and accumulates the THREE kh taps at once:
//   * the dY tile is loaded with its TMA box start at w0 = -(kw-1): the one-pixel shift and its zero
//     column come from TMA's out-of-bounds fill, so pixel rows never wrap across image rows;
//   * ONE X tile with a halo image row above and below serves all three kh: the tap's row shift
//     (kh * W pixels = kh * W * 128 B, a multiple of 1024 B for W >= 8) is applied to the start address
//     of the MN-major shared-memory descriptor (128B-swizzle phase is a function of the absolute
//     address: tools/experiments/umma_shift_test.cu);
//   * three fp32 accumulators [128 x kN] live side by side in TMEM (384 columns for kN = 128).
// Per 128-pixel stage this moves 68-80 KB through TMA for 24 MMAs (1536 tensor cycles) instead of 64 KB
// for 8 MMAs (512 cycles) in the one-tap-per-CTA version: 2.5x less L2->SMEM traffic per FLOP.
template <int kN> struct KwCfg {
  static constexpr int kBSlabs = kN / 64;
  static constexpr uint32_t kTmemCols = (3 * kN <= 256) ? 256 : 512;
};

struct KwParams {
  int B, H, W, Cin, Cout;
  int bh, tiles_h;
  int m_tiles, n_tiles, splits, kw_groups;
  int pair;                // Cout == 64: M rows 0..63 = kw_a, rows 64..127 = kw_b
  int x_rows;              // 128 + 2*W pixel rows in the X halo tile
  int stage_bytes, stages;
  int k_total;
  float* out;              // [splits][9][Cout][Cin]
  long long slab_stride;
};

template <int kN>
__global__ void __launch_bounds__(kThreadsW, 1)
conv3x3_wgrad_kw_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                        const KwParams p) {
  using C = KwCfg<kN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[4], empty_bar[4], done_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int rest = blockIdx.x;
  const int split = rest % p.splits; rest /= p.splits;
  const int n_tile = rest % p.n_tiles; rest /= p.n_tiles;
  const int m_tile = rest % p.m_tiles;
  const int kwg = rest / p.m_tiles;
  // kw of accumulator rows [0,64) and [64,128)
  const int kw_a = p.pair ? 2 * kwg : kwg;
  const int kw_b = p.pair ? (2 * kwg + 1 < 3 ? 2 * kwg + 1 : -1) : kwg;
  const int per = (p.k_total + p.splits - 1) / p.splits;
  const int k_begin = split * per;
  const int k_end = min(p.k_total, k_begin + per);
  const uint32_t x_slab_bytes = (uint32_t)p.x_rows * 128u;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int a_slabs = p.pair ? (kw_b >= 0 ? 2 : 1) : (p.Cout >= 128 ? 2 : 1);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)a_slabs * kSlabBytes + (uint32_t)C::kBSlabs * x_slab_bytes;
      for (int kt = k_begin; kt < k_end; ++kt) {
        const int b = kt / p.tiles_h;
        const int h0 = (kt % p.tiles_h) * p.bh;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * p.stage_bytes;
        uint8_t* sx = sa + 2 * kSlabBytes;
        mbar_arrive_expect_tx(&full_bar[stage], bytes);
        if (p.pair) {
          tma_load_4d(sa, &tmap_dy, &full_bar[stage], 0, -(kw_a - 1), h0, b);
          if (kw_b >= 0) tma_load_4d(sa + kSlabBytes, &tmap_dy, &full_bar[stage], 0, -(kw_b - 1), h0, b);
        } else {
          for (int s = 0; s < a_slabs; ++s)
            tma_load_4d(sa + s * kSlabBytes, &tmap_dy, &full_bar[stage], m_tile * 128 + s * 64, -(kw_a - 1), h0, b);
        }
#pragma unroll
        for (int s = 0; s < C::kBSlabs; ++s)
          tma_load_4d(sx + s * x_slab_bytes, &tmap_x, &full_bar[stage], n_tile * kN + s * 64, 0, h0 - 1, b);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kN, 1, 1);
      const uint32_t a_lbo = a_slabs == 2 ? kSlabBytes : 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k_begin; kt < k_end; ++kt) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t sx = sa + 2 * kSlabBytes;
        if (kN == 64) {
          // Cin = 64: the three kh taps are stacked in N.  The B operand is MN-major, i.e. consecutive 64-channel chunks
          // of N sit LBO bytes apart -- and the X tile of tap kh is the SAME halo tile one image row (W * 128 B) further
          // down.  With LBO = W * 128 one N = 192 instruction reads the dY slab ONCE for all three taps and fills the
          // same three accumulators (columns kh * 64 ...) as three N = 64 instructions would: a third of the A-operand
          // shared-memory reads.  (ncu r02: the N = 64 instructions kept the tensor pipe 56 % busy at 82 % L1/TEX.)
          constexpr uint32_t idesc3 = umma_idesc_bf16(128, 192, 1, 1);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                            // 128 pixels = 8 x K16
            const uint64_t da = umma_desc_sw128(sa + ks * 2048, a_lbo, 1024);
            const uint64_t db = umma_desc_sw128(sx + ks * 2048, (uint32_t)p.W * 128u, 1024);
            umma_bf16(tmem_base, da, db, idesc3, (kt > k_begin || ks > 0) ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t sxk = sx + (uint32_t)(kh * p.W) * 128u;    // halo tile starts at image row h0 - 1
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                          // 128 pixels = 8 x K16
              const uint64_t da = umma_desc_sw128(sa + ks * 2048, a_lbo, 1024);
              const uint64_t db = umma_desc_sw128(sxk + ks * 2048, x_slab_bytes, 1024);
              umma_bf16(tmem_base + kh * kN, da, db, idesc, (kt > k_begin || ks > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;                    // accumulator row
    int kw, co;
    bool valid;
    if (p.pair) {
      kw = row < 64 ? kw_a : kw_b;
      co = row & 63;
      valid = kw >= 0;
    } else {
      kw = kw_a;
      co = m_tile * 128 + row;
      valid = co < p.Cout;
    }
    const bool have = k_end > k_begin;
    if (have) {
      mbar_wait(&done_bar, 0);
      tcgen05_fence_after();
    }
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      const int tap = kh * 3 + (valid ? kw : 0);
      float* dst = p.out + (long long)split * p.slab_stride + ((long long)tap * p.Cout + (valid ? co : 0)) * p.Cin + n_tile * kN;
#pragma unroll 1
      for (int c = 0; c < kN / 32; ++c) {
        float v[32];
        if (have) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + kh * kN + c * 32, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (valid) {
          float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) d4[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// The kw-group weight gradient on CTA PAIRS (tcgen05.mma.cta_group::2) for Cout >= 256, Cin >= 128:
// one M = 256 MMA covers two 128-row Cout tiles (cluster rank r loads the dY slabs of tile 2*pair + r) and the
// N = 128 Cin tile is split between the CTAs: each one stages only ONE 64-channel X halo slab.  Per SM and K16 step
// that is 4 KB (A) + 2 KB (B) of operand reads instead of 4 + 4, and 52 KB instead of 72 KB through TMA per
// 128-pixel stage -- shared-memory bandwidth is what bounds the single-CTA kernel (ncu: TC pipe 83-87 % busy at
// ~60 % of the MMA peak).  Barrier protocol as in conv_halo2_tc.cu.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsW, 1)
conv3x3_wgrad_kw2_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                         const KwParams p) {
  constexpr int kN = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[4], empty_bar[4], done_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  int rest = blockIdx.x >> 1;                       // pair index
  const int split = rest % p.splits; rest /= p.splits;
  const int n_tile = rest % p.n_tiles; rest /= p.n_tiles;
  const int m_pair = rest % (p.m_tiles / 2);
  const int kw = rest / (p.m_tiles / 2);
  const int m_tile = 2 * m_pair + (int)rank;
  const int per = (p.k_total + p.splits - 1) / p.splits;
  const int k_begin = split * per;
  const int k_end = min(p.k_total, k_begin + per);
  const uint32_t x_slab_bytes = (uint32_t)p.x_rows * 128u;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = 2u * kSlabBytes + x_slab_bytes;     // per CTA
      for (int kt = k_begin; kt < k_end; ++kt) {
        const int b = kt / p.tiles_h;
        const int h0 = (kt % p.tiles_h) * p.bh;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * p.stage_bytes;
        uint8_t* sx = sa + 2 * kSlabBytes;
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * bytes);
        for (int s = 0; s < 2; ++s)
          tma_load_4d_2cta(sa + s * kSlabBytes, &tmap_dy, &full_bar[stage], m_tile * 128 + s * 64, -(kw - 1), h0, b);
        tma_load_4d_2cta(sx, &tmap_x, &full_bar[stage], n_tile * kN + (int)rank * 64, 0, h0 - 1, b);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, kN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = k_begin; kt < k_end; ++kt) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t sx = sa + 2 * kSlabBytes;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const uint32_t sxk = sx + (uint32_t)(kh * p.W) * 128u;      // halo tile starts at image row h0 - 1
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                            // 128 pixels = 8 x K16
            const uint64_t da = umma_desc_sw128(sa + ks * 2048, kSlabBytes, 1024);
            const uint64_t db = umma_desc_sw128(sxk + ks * 2048, x_slab_bytes, 1024);
            umma_bf16_2cta(tmem_base + kh * kN, da, db, idesc, (kt > k_begin || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit_2cta(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_2cta(&done_bar);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;                    // accumulator row = output channel within this CTA's tile
    const int co = m_tile * 128 + row;
    const bool have = k_end > k_begin;
    if (have) {
      mbar_wait(&done_bar, 0);
      tcgen05_fence_after();
    }
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      const int tap = kh * 3 + kw;
      float* dst = p.out + (long long)split * p.slab_stride + ((long long)tap * p.Cout + co) * p.Cin + n_tile * kN;
#pragma unroll 1
      for (int c = 0; c < kN / 32; ++c) {
        float v[32];
        if (have) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + kh * kN + c * 32, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
        for (int g = 0; g < 8; ++g) d4[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

struct KwPlan {
  int kN, kw_groups, m_tiles, n_tiles, splits, stages, stage_bytes, x_rows, pair, bh, tiles_h, k_total;
  int cta2;     // CTA-pair kernel (Cout >= 256, Cin >= 128): one X slab per CTA
};

bool g_wgrad_cta2 = true;
constexpr int kWgradSmemBudget = 208 * 1024;

KwPlan make_kw_plan(int B, int H, int W, int Cin, int Cout) {
  KwPlan q;
  q.kN = Cin >= 128 ? 128 : 64;
  q.pair = Cout == 64 ? 1 : 0;
  q.kw_groups = q.pair ? 2 : 3;
  q.m_tiles = q.pair ? 1 : (Cout + 127) / 128;
  q.n_tiles = Cin / q.kN;
  q.bh = 128 / W;
  q.tiles_h = (H + q.bh - 1) / q.bh;
  q.k_total = B * q.tiles_h;
  q.x_rows = 128 + 2 * W;
  q.cta2 = (g_wgrad_cta2 && Cout % 256 == 0 && q.kN == 128) ? 1 : 0;
  q.stage_bytes = 2 * kSlabBytes + (q.cta2 ? 1 : q.kN / 64) * q.x_rows * 128;
  // Shared-memory budget: 208 KB, not the full 227 KB.  The weight gradient of layer l runs on a side stream while the
  // HBM-bound BatchNorm backward of layer l-1 runs on the main one (engine.trunk_backward); three of those 256-thread
  // CTAs (5 KB each incl. the per-CTA reserve) must fit on the SM beside this CTA for the two to overlap.  Only the
  // 64 -> 128 layer loses a stage (4 -> 3) to this.
  q.stages = kWgradSmemBudget / q.stage_bytes;
  if (q.stages > 4) q.stages = 4;
  // One CTA per SM fits (shared memory), so the grid must not exceed ONE wave: with ceil(2 * SMs / items) splits
  // most layers launched 297-336 CTAs on 148 SMs, i.e. a nearly empty third wave (+50 % time).
  const int items = q.kw_groups * q.m_tiles * q.n_tiles;               // CTAs per split (pairs count as 2 CTAs)
  long long splits = sm_count() / items;
  if (splits > q.k_total) splits = q.k_total;
  if (splits < 1) splits = 1;
  q.splits = (int)splits;
  return q;
}

template <int kN>
int launch_wgrad_kw(const CUtensorMap& tdy, const CUtensorMap& tx, const KwParams& p, int grid, cudaStream_t stream) {
  auto kern = conv3x3_wgrad_kw_kernel<kN>;
  const int smem = p.stages * p.stage_bytes + 1024;
  SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, kThreadsW, smem, stream>>>(tdy, tx, p);
  SED_LAUNCH_CHECK("conv3x3_wgrad_kw_kernel");
  return 0;
}

template <int kN>
int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgradParams& p, int grid, cudaStream_t stream) {
  auto kern = conv3x3_wgrad_tc_kernel<kN>;
  SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WCfg<kN>::kDynSmem));
  kern<<<grid, kThreadsW, WCfg<kN>::kDynSmem, stream>>>(tdy, tx, p);
  SED_LAUNCH_CHECK("conv3x3_wgrad_tc_kernel");
  return 0;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

// Development switch (tests / A-B timing): 0 = single-CTA weight-gradient kernel everywhere, 1 = CTA pairs where they apply.
int sed_conv3x3_tc_wgrad_use_pairs(int on) {
  const int was = g_wgrad_cta2 ? 1 : 0;
  g_wgrad_cta2 = on != 0;
  return was;
}

// Number of split-K slabs sed_conv3x3_tc_wgrad will write for this shape (workspace sizing).
int sed_conv3x3_tc_wgrad_splits(int B, int H, int W, int Cin, int Cout) {
  if (W < 8 || 128 % W != 0 || Cin < 64 || Cout < 64) return 0;
  return make_kw_plan(B, H, W, Cin, Cout).splits;
}

int sed_conv3x3_tc_wgrad(const void* dy, const void* x, float* slabs, int B, int H, int W, int Cin, int Cout,
                         sed_stream_t stream) {
  SED_REQUIRE(dy && x && slabs, "sed_conv3x3_tc_wgrad: null pointer");
  SED_REQUIRE(B > 0 && H > 0, "sed_conv3x3_tc_wgrad: empty batch");
  SED_REQUIRE(W >= 8 && W <= 64 && 128 % W == 0, "sed_conv3x3_tc_wgrad: W=%d must divide 128 and lie in [8, 64]", W);
  SED_REQUIRE(Cin % 64 == 0 && Cin >= 64 && (Cin == 64 || Cin % 128 == 0), "sed_conv3x3_tc_wgrad: Cin=%d unsupported", Cin);
  SED_REQUIRE(Cout == 64 || Cout % 128 == 0, "sed_conv3x3_tc_wgrad: Cout=%d unsupported", Cout);
  SED_REQUIRE(aligned(slabs, 16), "sed_conv3x3_tc_wgrad: workspace must be 16-byte aligned");
  const KwPlan q = make_kw_plan(B, H, W, Cin, Cout);
  SED_REQUIRE(q.stages >= 2, "sed_conv3x3_tc_wgrad: stage of %d bytes does not fit twice in shared memory", q.stage_bytes);
  KwParams p;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bh = q.bh; p.tiles_h = q.tiles_h;
  p.m_tiles = q.m_tiles; p.n_tiles = q.n_tiles; p.splits = q.splits; p.kw_groups = q.kw_groups;
  p.pair = q.pair; p.x_rows = q.x_rows; p.stage_bytes = q.stage_bytes; p.stages = q.stages;
  p.k_total = q.k_total;
  p.out = slabs;
  p.slab_stride = 9LL * Cout * Cin;
  const int grid = q.kw_groups * q.m_tiles * q.n_tiles * q.splits;

  alignas(64) CUtensorMap tdy, tx;
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)q.bh, 1};
    if (int rc = tc::make_tmap_bf16(&tdy, dy, 4, dims, strides, box, "wgrad dY map")) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)(q.bh + 2), 1};
    if (int rc = tc::make_tmap_bf16(&tx, x, 4, dims, strides, box, "wgrad X halo map")) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (q.cta2) {
    const int smem = p.stages * p.stage_bytes + 1024;
    SED_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kw2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    conv3x3_wgrad_kw2_kernel<<<grid, kThreadsW, smem, s>>>(tdy, tx, p);      // grid = pairs x 2: m_tiles is even here
    SED_LAUNCH_CHECK("conv3x3_wgrad_kw2_kernel");
    return 0;
  }
  if (q.kN == 128) return launch_wgrad_kw<128>(tdy, tx, p, grid, s);
  return launch_wgrad_kw<64>(tdy, tx, p, grid, s);
}


// out[M][N] = A[R][M]^T * Bm[R][N]  (A, Bm bf16 row-major, i.e. both "MN-major" with the reduction
// index R strided): the weight-gradient GEMMs of the GRU / attention projections
// (dW = dOut^T X).  Split-K slabs [splits][M][N] fp32; reduce with sed_reduce_partials.
int sed_gemm_tn_tc_splits(long long R, int M, int N) {
  const int kN = N >= 128 ? 128 : 64;
  const int items = ((M + 127) / 128) * (N / kN);
  const long long k_total = (R + 127) / 128;
  long long splits = sm_count() / items;                         // one wave (one CTA fits per SM)
  if (splits > k_total) splits = k_total;
  if (splits < 1) splits = 1;
  return (int)splits;
}

int sed_gemm_tn_tc(const void* a, int lda, const void* bm, int ldb, float* slabs, long long R, int M, int N,
                   sed_stream_t stream) {
  SED_REQUIRE(a && bm && slabs, "sed_gemm_tn_tc: null pointer");
  SED_REQUIRE(R > 0 && R < (1LL << 31), "sed_gemm_tn_tc: bad R");
  SED_REQUIRE(M == 64 || M % 128 == 0, "sed_gemm_tn_tc: M=%d unsupported", M);
  SED_REQUIRE(N == 64 || N % 128 == 0, "sed_gemm_tn_tc: N=%d unsupported", N);
  SED_REQUIRE(lda >= M && ldb >= N && lda % 8 == 0 && ldb % 8 == 0, "sed_gemm_tn_tc: bad leading dimensions");
  const int kN = N >= 128 ? 128 : 64;
  WgradParams p;
  p.B = 1; p.H = 1; p.W = 128; p.Cin = N; p.Cout = M;
  p.bh = 1; p.tiles_h = (int)((R + 127) / 128);
  p.m_tiles = (M + 127) / 128;
  p.n_tiles = N / kN;
  p.splits = sed_gemm_tn_tc_splits(R, M, N);
  p.a_slabs = M == 64 ? 1 : 2;
  p.gemm = 1;
  p.k_total = p.tiles_h;
  p.out = slabs;
  p.slab_stride = (long long)M * N;
  const int grid = p.m_tiles * p.n_tiles * p.splits;       // tap index is always 0
  alignas(64) CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)M, (uint64_t)R, 1, 1};
    const uint64_t strides[3] = {(uint64_t)lda * 2, (uint64_t)R * lda * 2, (uint64_t)R * lda * 2};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int rc = tc::make_tmap_bf16(&ta, a, 4, dims, strides, box, "gemm_tn A map")) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)R, 1, 1};
    const uint64_t strides[3] = {(uint64_t)ldb * 2, (uint64_t)R * ldb * 2, (uint64_t)R * ldb * 2};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int rc = tc::make_tmap_bf16(&tb, bm, 4, dims, strides, box, "gemm_tn B map")) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (kN == 128) return launch_wgrad<128>(ta, tb, p, grid, s);
  return launch_wgrad<64>(ta, tb, p, grid, s);
}

}  // extern "C"
