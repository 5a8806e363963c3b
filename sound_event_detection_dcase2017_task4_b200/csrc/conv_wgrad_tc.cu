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

// Number of split-K slabs sed_conv3x3_tc_wgrad will write for this shape (workspace sizing).
int sed_conv3x3_tc_wgrad_splits(int B, int H, int W, int Cin, int Cout) {
  if (W <= 0 || 128 % W != 0) return 0;
  const int bh = 128 / W;
  const int kN = Cin >= 128 ? 128 : 64;
  const int items = 9 * ((Cout + 127) / 128) * (Cin / kN);
  const long long k_total = (long long)B * ((H + bh - 1) / bh);
  long long splits = (2LL * sm_count() + items - 1) / items;      // ~2 CTAs worth of work per SM
  if (splits > k_total) splits = k_total;
  if (splits < 1) splits = 1;
  return (int)splits;
}

int sed_conv3x3_tc_wgrad(const void* dy, const void* x, float* slabs, int B, int H, int W, int Cin, int Cout,
                         sed_stream_t stream) {
  SED_REQUIRE(dy && x && slabs, "sed_conv3x3_tc_wgrad: null pointer");
  SED_REQUIRE(B > 0 && H > 0, "sed_conv3x3_tc_wgrad: empty batch");
  SED_REQUIRE(W >= 8 && W <= 128 && 128 % W == 0, "sed_conv3x3_tc_wgrad: W=%d must divide 128 and be >= 8", W);
  SED_REQUIRE(Cin % 64 == 0 && Cin >= 64 && (Cin == 64 || Cin % 128 == 0), "sed_conv3x3_tc_wgrad: Cin=%d unsupported", Cin);
  SED_REQUIRE(Cout == 64 || Cout % 128 == 0, "sed_conv3x3_tc_wgrad: Cout=%d unsupported", Cout);
  SED_REQUIRE(aligned(slabs, 16), "sed_conv3x3_tc_wgrad: workspace must be 16-byte aligned");
  const int kN = Cin >= 128 ? 128 : 64;
  WgradParams p;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bh = 128 / W;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.m_tiles = (Cout + 127) / 128;
  p.n_tiles = Cin / kN;
  p.splits = sed_conv3x3_tc_wgrad_splits(B, H, W, Cin, Cout);
  p.a_slabs = Cout == 64 ? 1 : 2;
  p.gemm = 0;
  p.k_total = B * p.tiles_h;
  p.out = slabs;
  p.slab_stride = 9LL * Cout * Cin;
  const int grid = 9 * p.m_tiles * p.n_tiles * p.splits;

  alignas(64) CUtensorMap tdy, tx;
  {
    const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)p.bh, 1};
    if (int rc = tc::make_tmap_bf16(&tdy, dy, 4, dims, strides, box, "wgrad dY map")) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)p.bh, 1};
    if (int rc = tc::make_tmap_bf16(&tx, x, 4, dims, strides, box, "wgrad X map")) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (kN == 128) return launch_wgrad<128>(tdy, tx, p, grid, s);
  return launch_wgrad<64>(tdy, tx, p, grid, s);
}


// out[M][N] = A[R][M]^T * Bm[R][N]  (A, Bm bf16 row-major, i.e. both "MN-major" with the reduction
// index R strided): the weight-gradient GEMMs of the GRU / attention projections
// (dW = dOut^T X).  Split-K slabs [splits][M][N] fp32; reduce with sed_reduce_partials.
int sed_gemm_tn_tc_splits(long long R, int M, int N) {
  const int kN = N >= 128 ? 128 : 64;
  const int items = ((M + 127) / 128) * (N / kN);
  const long long k_total = (R + 127) / 128;
  long long splits = (2LL * sm_count() + items - 1) / items;
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
