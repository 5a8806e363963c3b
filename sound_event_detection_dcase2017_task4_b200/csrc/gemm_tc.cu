// gemm_tc.cu -- plain tensor-core GEMM on the one-box-per-K-block tcgen05 pipeline (sm_100a):
//   out[M][N] = A[M][K] * Bw[N][K]^T (+ bias),  A / Bw bf16 K-major, out fp32.
// Used for the GRU / attention projections (x @ W^T) of /root/reference/pytorch/models.py:475/:566 (nn.GRU input
// projection) and :641-665 (MultiHead), optionally with the 3-way bf16 operand split (sed_split_bf16x3) for
// fp32-class accuracy.  (The 3x3 convolutions moved to the halo-tile kernel in conv_halo_tc.cu.)
//
// warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-9 = epilogue; 128B-swizzled K-major tiles
// consumed directly by tcgen05.mma.kind::f16 (M = 128, N tile <= 256, K = 16), fp32 accumulators double buffered in
// TMEM, persistent static round-robin tile schedule.
#include "common.cuh"
#include "tc.cuh"

namespace sed {
namespace {

using namespace tc;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = one 128 B swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kNumThreads = 320;                  // 10 warps: TMA, MMA, 8 epilogue
constexpr int kEpiThreads = 256;                  // two warps per TMEM lane quarter, interleaved over 32-column chunks

template <int kN> struct Cfg {
  static constexpr int kBBytes = kN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kN == 256) ? 4 : (kN == 128 ? 6 : 8);
  static constexpr uint32_t kTmemCols = (2 * kN <= 32) ? 32 : (2 * kN <= 64) ? 64 : (2 * kN <= 128) ? 128 : (2 * kN <= 256) ? 256 : 512;
  static constexpr int kDynSmem = kStages * kStageBytes + 1024;   // +1024 for manual alignment
};

struct GemmParams {
  int N;             // columns of the output
  int tiles_m;       // ceil(M / 128)
  int tiles_n;       // N / kN
  int num_tiles;     // tiles_m * tiles_n
  int gemm_m, gemm_kb;        // rows of A, K / 64
  float* out_f32;    // D[M][N] = A[M][K] * B[N][K]^T (+ bias[N]), fp32 row-major
  const float* bias;
};

template <int kN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmParams p) {
  using C = Cfg<kN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024 B aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[C::kStages], empty_bar[C::kStages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.gemm_kb;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        const int m_tile = tile / p.tiles_n;
        {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * C::kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
            tma_load_4d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_tile * kBlockM, 0, 0);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBlockK, n_tile * kN);
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int ks = 0; ks < kBlockK / 16; ++ks) {
            const uint64_t da = umma_desc_sw128(sa + ks * 32, 16, 1024);
            const uint64_t db = umma_desc_sw128(sb + ks * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);          // smem slot reusable once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);          // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;              // which of the two warps of this quarter
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile = tile % p.tiles_n;
      const int m_tile = tile / p.tiles_n;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      {
        const long long m = (long long)m_tile * kBlockM + row;
        const bool ok = m < p.gemm_m;
        float* drow = p.out_f32 + m * p.N + n_tile * kN;
#pragma unroll 1
        for (int c = half; c < kN / 32; c += 2) {
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kN + c * 32, v);
          if (ok) {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + n_tile * kN + c * 32 + i);
            }
            float4* d4 = reinterpret_cast<float4*>(drow + c * 32);
#pragma unroll
            for (int g = 0; g < 8; ++g) d4[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
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

template <int kN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = gemm_nt_tc_kernel<kN>;
  SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<kN>::kDynSmem));
  kern<<<grid, kNumThreads, Cfg<kN>::kDynSmem, stream>>>(ta, tb, p);
  SED_LAUNCH_CHECK("gemm_nt_tc_kernel");
  return 0;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

// Plain tensor-core GEMM on the same pipeline: out[M][N] = A[M][K] * Bw[N][K]^T (+ bias), A / Bw bf16
// K-major, out fp32.  K % 64 == 0, N in {64, 128} or a multiple of 256.  Used for the GRU / attention
// projections (x @ W^T), optionally with the 3-way bf16 split (see sed_split_bf16x3) for fp32-class accuracy.
int sed_gemm_tc(const void* a, const void* bw, const float* bias, float* out, long long M, int N, int K,
                sed_stream_t stream) {
  SED_REQUIRE(a && bw && out, "sed_gemm_tc: null pointer");
  SED_REQUIRE(M > 0 && M < (1LL << 31), "sed_gemm_tc: bad M");
  SED_REQUIRE(K % 64 == 0 && K >= 64, "sed_gemm_tc: K=%d must be a multiple of 64", K);
  SED_REQUIRE(N == 64 || N == 128 || N % 256 == 0, "sed_gemm_tc: N=%d unsupported", N);
  SED_REQUIRE(aligned(out, 16), "sed_gemm_tc: output must be 16-byte aligned");
  const int kN = N >= 256 ? 256 : N;
  GemmParams p;
  p.N = N;
  p.tiles_m = (int)((M + 127) / 128);
  p.tiles_n = N / kN;
  p.num_tiles = p.tiles_m * p.tiles_n;
  p.gemm_m = (int)M; p.gemm_kb = K / 64; p.out_f32 = out; p.bias = bias;
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  alignas(64) CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, 1, 1};
    const uint64_t strides[3] = {(uint64_t)K * 2, (uint64_t)M * K * 2, (uint64_t)M * K * 2};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int rc = tc::make_tmap_bf16(&ta, a, 4, dims, strides, box, "gemm A map")) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)kN};
    if (int rc = tc::make_tmap_bf16(&tb, bw, 2, dims, strides, box, "gemm B map")) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (kN) {
    case 64: return launch_gemm<64>(ta, tb, p, grid, s);
    case 128: return launch_gemm<128>(ta, tb, p, grid, s);
    case 256: return launch_gemm<256>(ta, tb, p, grid, s);
  }
  SED_REQUIRE(false, "sed_gemm_tc: no kernel for N tile %d", kN);
}

}  // extern "C"
