// conv_halo_tc.cu -- 3x3 / stride 1 / pad 1 convolution as a tcgen05 implicit GEMM (sm_100a), halo-tile version.
//
// Replaces the cuDNN calls behind ConvBlock.conv1/conv2 (/root/reference/pytorch/models.py:75-83,
// forward :102-103) for every layer with Cin % 64 == 0, and -- with the rotated / transposed weight
// pack -- their data gradient (loss.backward(), /root/reference/pytorch/main.py:257).
//
//   D[mt][128 pixels][kN] = sum over kw, 64-channel blocks cb, kh   A(kw,cb)[rows shifted by kh][64] * W(kh,kw,cb)[kN][64]^T
//
// Why a halo tile.  With one TMA box per tap every activation byte crosses L2 -> shared memory nine times and
// every weight byte once per 128 pixels; at 1.2+ PFLOP/s that traffic (13-15 TB/s measured over the whole
// chip) saturates the L2 -> SM path before the tensor pipe.  Here a CTA owns kMT vertically adjacent M tiles
// (kMT * bh image rows, bh * W = 128 pixels each) and, per (kw, cb), loads ONE activation tile that carries a
// halo image row above and below: a 4-D TMA box {64 ch, W, kMT*bh + 2, 1} at {cb*64, kw-1, h0-1, b}.
//   * the three kh taps of every M tile read that same tile through shared-memory descriptors whose start
//     address is advanced by whole image rows ((kh + mt*bh) * W * 128 B, a multiple of the 1024 B swizzle
//     atom for W >= 8, so the 128B-swizzle phase is unchanged);
//   * the kw shift and all zero padding (left / right columns, rows -1 and H, the H tail) come from TMA's
//     out-of-bounds zero fill, so an im2col matrix never exists anywhere;
//   * each weight tile W(kh,kw,cb) (kN x 64, K-major) is loaded once per CTA tile and used by all kMT M tiles.
// L2 -> shared-memory bytes per 128 pixels x kN outputs x 64 input channels x 3 kh taps:
//   before: 3 * (16 KB + kN*128 B);   now: (kMT*bh+2)/(kMT*bh) * 16 KB + 3*kN*128 B / kMT.
// For Cin = 64 and kN = 64 (block 1) the whole 9-tap weight set (72 KB) stays resident in shared memory.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-9 = epilogue (two warps per
// TMEM lane quarter, interleaved over 32-column chunks): tcgen05.ld -> bf16 NHWC store, plus per-channel sum /
// sum of squares of the fp32 accumulators for the training-mode BatchNorm that follows (warp-transpose
// reduction, accumulated per CTA in shared memory, flushed once: deterministic two-level reduction).
// fp32 accumulators are double buffered in TMEM (2 x kMT x kN columns): the epilogue of tile i overlaps the
// MMAs of tile i+1.  Persistent: grid = min(#tiles, #SMs), static round-robin tile schedule.
#include "common.cuh"
#include "tc.cuh"

namespace sed {
namespace {

using namespace tc;

constexpr int kHThreads = 320;          // 10 warps: TMA, MMA, 8 epilogue
constexpr int kHEpiThreads = 256;
constexpr int kMaxAStages = 4;
constexpr int kMaxBStages = 8;
constexpr int kSmemBudget = 220 * 1024; // dynamic shared memory the stage rings may use (227 KB - static - alignment)

// 32x32 warp transpose-reduce of two register arrays: afterwards lane l holds in a[0] / b[0] the total of column l
// over the warp's 32 rows (fixed order: deterministic).
__device__ __forceinline__ void warp_transpose_reduce2(float (&a)[32], float (&b)[32], int lane) {
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

struct HaloParams {
  int B, H, W, Cin, Cout;
  int bh;              // image rows per M tile; bh * W == 128
  int rows_per_tile;   // kMT * bh
  int tiles_h;         // ceil(H / rows_per_tile)
  int tiles_n;         // Cout / kN
  int num_tiles;       // B * tiles_h * tiles_n
  int kb_per_tap;      // Cin / 64
  int a_bytes;         // (rows_per_tile + 2) * W * 128
  int a_stages, b_stages;
  __nv_bfloat16* y;    // NHWC output
  float* stats;        // [gridDim.x * 4][2][Cout] partial (sum, sumsq) rows, one per (CTA, TMEM lane quarter), or nullptr
  // fused first pass of the BatchNorm+ReLU(+2x2 avg-pool) backward of the layer BELOW (kBnr kernels, data gradient):
  // this kernel's output is that layer's dA; partials of sum(g) and sum(g*y), g = unpool(dA)/pool^2 * [bn(y) > 0]
  const __nv_bfloat16* bnr_y;      // raw conv output of the layer below: [B][H*pool (+tail)][W*pool][Cout]
  const float* bnr_scale;          // its BatchNorm scale / shift (relu mask = y*scale + shift > 0)
  const float* bnr_shift;
  int bnr_pool;                    // 1 or 2
  int bnr_Hy;                      // rows of bnr_y per clip
};

__device__ __forceinline__ void unpack8_bf16(const uint4& raw, float (&v)[8]) {
  const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

template <int kN, int kMT, bool kResidentB, bool kBnr>
__global__ void __launch_bounds__(kHThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const HaloParams p) {
  constexpr int kBBytes = kN * 128;                       // one tap, one 64-channel block
  constexpr uint32_t kTmemCols = 2 * kMT * kN;
  static_assert(kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns must be a power of two <= 512");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.a_stages * p.a_bytes;
  __shared__ uint64_t a_full[kMaxAStages], a_empty[kMaxAStages], b_full[kMaxBStages], b_empty[kMaxBStages];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2], b_res_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kMaxAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], kHEpiThreads); }
    mbar_init(&b_res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (kResidentB) {
        const int nb = 9 * p.kb_per_tap;
        mbar_arrive_expect_tx(&b_res_full, (uint32_t)(nb * kBBytes));
        for (int i = 0; i < nb; ++i)                       // slot i = tap * kb_per_tap + cb  <->  k offset i * 64
          tma_load_2d(smem_b + i * kBBytes, &tmap_b, &b_res_full, i * 64, 0);
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        const int m_tile = tile / p.tiles_n;
        const int b = m_tile / p.tiles_h;
        const int h0 = (m_tile % p.tiles_h) * p.rows_per_tile;
        for (int kw = 0; kw < 3; ++kw) {
          for (int cb = 0; cb < p.kb_per_tap; ++cb) {
            mbar_wait(&a_empty[as], aph ^ 1);
            mbar_arrive_expect_tx(&a_full[as], (uint32_t)p.a_bytes);
            tma_load_4d(smem_a + as * p.a_bytes, &tmap_a, &a_full[as], cb * 64, kw - 1, h0 - 1, b);
            if (++as == p.a_stages) { as = 0; aph ^= 1; }
            if (!kResidentB) {
#pragma unroll 1
              for (int kh = 0; kh < 3; ++kh) {
                mbar_wait(&b_empty[bs], bph ^ 1);
                mbar_arrive_expect_tx(&b_full[bs], (uint32_t)kBBytes);
                tma_load_2d(smem_b + bs * kBBytes, &tmap_b, &b_full[bs], (kh * 3 + kw) * p.Cin + cb * 64, n_tile * kN);
                if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kN, 0, 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      if (kResidentB) {
        mbar_wait(&b_res_full, 0);
        tcgen05_fence_after();
      }
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (kMT * kN);
        for (int kw = 0; kw < 3; ++kw) {
          for (int cb = 0; cb < p.kb_per_tap; ++cb) {
            mbar_wait(&a_full[as], aph);
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem_a + as * p.a_bytes);
#pragma unroll 1
            for (int kh = 0; kh < 3; ++kh) {
              uint32_t sb;
              if (kResidentB) {
                sb = smem_u32(smem_b + ((kh * 3 + kw) * p.kb_per_tap + cb) * kBBytes);
              } else {
                mbar_wait(&b_full[bs], bph);
                tcgen05_fence_after();
                sb = smem_u32(smem_b + bs * kBBytes);
              }
              const uint32_t first = (kw | cb | kh) == 0 ? 0u : 1u;
#pragma unroll
              for (int mt = 0; mt < kMT; ++mt) {
                const uint32_t sam = sa + (uint32_t)((kh + mt * p.bh) * p.W) * 128u;   // whole image rows: 1024 B multiples
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t da = umma_desc_sw128(sam + ks * 32, 16, 1024);
                  const uint64_t db = umma_desc_sw128(sb + ks * 32, 16, 1024);
                  umma_bf16(d_tmem + mt * kN, da, db, idesc, (first | (uint32_t)ks) != 0 ? 1u : 0u);
                }
              }
              if (!kResidentB) {
                umma_commit(&b_empty[bs]);               // weight slot reusable once these MMAs retire
                if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
              }
            }
            umma_commit(&a_empty[as]);                   // activation slot reusable
            if (++as == p.a_stages) { as = 0; aph ^= 1; }
          }
        }
        umma_commit(&tmem_full_bar[acc]);                // accumulators complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // (setmaxnreg register re-allocation over 3 warpgroups was tried: ptxas then spills in every role -- kept at 10 warps)
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;              // which of the two warps of this quarter
    const int row = q * 32 + lane;
    const int r_h = row / p.W, r_w = row % p.W;
    // BatchNorm statistics: every thread sums the fp32 accumulators of ITS pixel rows per column in registers --
    // over the kMT M tiles of a CTA tile, and for kN == 64 (one 32-column chunk per warp, K = 576: the layers whose
    // epilogue is on the critical path) over ALL tiles of the CTA -- before the 32x32 warp transpose-reduce runs.
    constexpr bool kLocalStats = kN == 64;
    constexpr int kChunksPerWarp = kN / 64;
    // Each (CTA, lane quarter) owns one partial row in global memory and each of its two warps owns every other
    // 32-channel chunk of it, so the running sums are plain read-modify-writes by their only writer: no atomics, and
    // the finalize kernel adds the rows in a fixed order -- bit-reproducible statistics.
    float* stat_row = p.stats == nullptr ? nullptr : p.stats + ((long long)blockIdx.x * 4 + q) * 2 * p.Cout;
    if (stat_row != nullptr && !kLocalStats) {
      for (int nt = 0; nt < p.tiles_n; ++nt)
#pragma unroll
        for (int j = 0; j < kChunksPerWarp; ++j) {
          const int ch = nt * kN + (half + 2 * j) * 32 + lane;
          stat_row[ch] = 0.f;
          stat_row[p.Cout + ch] = 0.f;
        }
    }
    float s_acc[32], ss_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s_acc[i] = ss_acc[i] = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_tile = tile % p.tiles_n;
      const int m_tile = tile / p.tiles_n;
      const int b = m_tile / p.tiles_h;
      const int h0 = (m_tile % p.tiles_h) * p.rows_per_tile;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll
      for (int j = 0; j < kChunksPerWarp; ++j) {
        const int c = half + 2 * j;
#pragma unroll 1
        for (int mt = 0; mt < kMT; ++mt) {
          const int h = h0 + mt * p.bh + r_h;
          const bool valid = h < p.H;
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (acc * kMT + mt) * kN + c * 32, v);
          if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(p.y + (((long long)b * p.H + h) * p.W + r_w) * p.Cout + n_tile * kN + c * 32);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#pragma unroll
            for (int g = 0; g < 4; ++g) d4[g] = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
            if (kBnr) {
              // g uses the bf16-rounded dA that was just stored (what the second BN-backward pass will read back)
              const int ch0 = n_tile * kN + c * 32;
              const float4* sc4 = reinterpret_cast<const float4*>(p.bnr_scale + ch0);
              const float4* sh4 = reinterpret_cast<const float4*>(p.bnr_shift + ch0);
              const int pool = p.bnr_pool;
              const float gscale = pool == 2 ? 0.25f : 1.0f;
              const int Wy = p.W * pool;
#pragma unroll 1
              for (int win = 0; win < pool * pool; ++win) {
                const int dh = win >> 1, dw = win & 1;
                const uint4* yp = reinterpret_cast<const uint4*>(
                    p.bnr_y + (((long long)b * p.bnr_Hy + h * pool + dh) * Wy + r_w * pool + dw) * p.Cout + ch0);
                uint4 raw[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) raw[g] = __ldg(yp + g);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  float yv[8];
                  unpack8_bf16(raw[g], yv);
                  const float4 sa = __ldg(sc4 + 2 * g), sb = __ldg(sc4 + 2 * g + 1);
                  const float4 ha = __ldg(sh4 + 2 * g), hb = __ldg(sh4 + 2 * g + 1);
                  const float scv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                  const float shv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float2 d2 = unpack_bf16x2(pk[4 * g + (k >> 1)]);
                    const float dv = ((k & 1) ? d2.y : d2.x) * gscale;
                    const float gk = fmaf(yv[k], scv[k], shv[k]) > 0.f ? dv : 0.f;
                    s_acc[8 * g + k] += gk;
                    ss_acc[8 * g + k] = fmaf(gk, yv[k], ss_acc[8 * g + k]);
                  }
                }
              }
            } else if (p.stats != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                s_acc[i] += v[i];
                ss_acc[i] = fmaf(v[i], v[i], ss_acc[i]);
              }
            }
          }
        }
        if (!kLocalStats && p.stats != nullptr) {
          warp_transpose_reduce2(s_acc, ss_acc, lane);
          const int ch = n_tile * kN + c * 32 + lane;
          stat_row[ch] += s_acc[0];
          stat_row[p.Cout + ch] += ss_acc[0];
#pragma unroll
          for (int i = 0; i < 32; ++i) s_acc[i] = ss_acc[i] = 0.f;
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (kLocalStats && stat_row != nullptr) {          // kN == 64: Cout == 64, a single n tile; one reduce per warp
      warp_transpose_reduce2(s_acc, ss_acc, lane);
      stat_row[half * 32 + lane] = s_acc[0];
      stat_row[p.Cout + half * 32 + lane] = ss_acc[0];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Tiling plan shared by the launcher and the workspace-size query.
struct HaloPlan {
  int kN, kMT, resident, bh, rows_per_tile, tiles_h, tiles_n, num_tiles, a_bytes, a_stages, b_stages, smem, grid;
};

bool make_plan(int B, int H, int W, int Cin, int Cout, HaloPlan* q) {
  if (W < 8 || W > 128 || 128 % W != 0 || Cin < 64 || Cin % 64 != 0) return false;
  if (!(Cout == 64 || Cout == 128 || Cout % 256 == 0) || Cout > 512) return false;
  q->kN = Cout == 64 ? 64 : (Cout == 128 ? 128 : 256);      // measured: N = 256 tiles beat 2 x (N = 128) for Cout >= 256
  q->kMT = q->kN == 256 ? 1 : 2;
  q->resident = (Cout == 64 && Cin == 64) ? 1 : 0;
  q->bh = 128 / W;
  q->rows_per_tile = q->kMT * q->bh;
  q->tiles_h = (H + q->rows_per_tile - 1) / q->rows_per_tile;
  q->tiles_n = Cout / q->kN;
  const long long tiles = (long long)B * q->tiles_h * q->tiles_n;
  if (tiles <= 0 || tiles >= (1LL << 31)) return false;
  q->num_tiles = (int)tiles;
  q->a_bytes = (q->rows_per_tile + 2) * W * 128;
  const int b_bytes = q->kN * 128;
  if (q->resident) {
    const int res = 9 * (Cin / 64) * b_bytes;
    q->a_stages = (kSmemBudget - res) / q->a_bytes;
    if (q->a_stages > kMaxAStages) q->a_stages = kMaxAStages;
    q->b_stages = 0;
    q->smem = res + q->a_stages * q->a_bytes + 1024;
  } else {
    q->a_stages = 3;
    q->b_stages = (kSmemBudget - q->a_stages * q->a_bytes) / b_bytes;
    if (q->b_stages > kMaxBStages) q->b_stages = kMaxBStages;
    q->smem = q->a_stages * q->a_bytes + q->b_stages * b_bytes + 1024;
  }
  if (q->a_stages < 2 || (!q->resident && q->b_stages < 3)) return false;
  q->grid = q->num_tiles < sm_count() ? q->num_tiles : sm_count();
  return true;
}

template <int kN, int kMT, bool kRes, bool kBnr>
int launch_halo(const CUtensorMap& ta, const CUtensorMap& tb, const HaloParams& p, const HaloPlan& q, cudaStream_t stream) {
  auto kern = conv3x3_halo_kernel<kN, kMT, kRes, kBnr>;
  SED_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, q.smem));
  kern<<<q.grid, kHThreads, q.smem, stream>>>(ta, tb, p);
  SED_LAUNCH_CHECK("conv3x3_halo_kernel");
  return 0;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_conv3x3_tc_grid(int B, int H, int W, int Cin, int Cout) {   // rows of the statistics workspace: 4 per CTA
  HaloPlan q;
  if (!make_plan(B, H, W, Cin, Cout, &q)) return 0;
  return q.grid * 4;
}

static int conv_halo_dispatch(const void* x, const void* wpack, void* y, float* stats_partial, int B, int H, int W,
                              int Cin, int Cout, const void* bnr_y, const float* bnr_scale, const float* bnr_shift,
                              int bnr_pool, int bnr_Hy, sed_stream_t stream, const char* who) {
  SED_REQUIRE(x && wpack && y, "%s: null pointer", who);
  SED_REQUIRE(B > 0 && H > 0, "%s: empty batch", who);
  SED_REQUIRE(W >= 8 && W <= 128 && 128 % W == 0, "%s: W=%d must divide 128 and be >= 8", who, W);
  SED_REQUIRE(Cin % 64 == 0 && Cin >= 64, "%s: Cin=%d must be a multiple of 64", who, Cin);
  SED_REQUIRE((Cout == 64 || Cout == 128 || Cout % 256 == 0) && Cout <= 512, "%s: Cout=%d unsupported", who, Cout);
  SED_REQUIRE(aligned(y, 16), "%s: output must be 16-byte aligned", who);
  HaloPlan q;
  SED_REQUIRE(make_plan(B, H, W, Cin, Cout, &q), "%s: no tiling for B=%d H=%d W=%d Cin=%d Cout=%d", who, B, H, W, Cin, Cout);
  HaloParams p;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bh = q.bh; p.rows_per_tile = q.rows_per_tile; p.tiles_h = q.tiles_h; p.tiles_n = q.tiles_n;
  p.num_tiles = q.num_tiles; p.kb_per_tap = Cin / 64; p.a_bytes = q.a_bytes;
  p.a_stages = q.a_stages; p.b_stages = q.b_stages;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.stats = stats_partial;
  p.bnr_y = reinterpret_cast<const __nv_bfloat16*>(bnr_y);
  p.bnr_scale = bnr_scale; p.bnr_shift = bnr_shift; p.bnr_pool = bnr_pool; p.bnr_Hy = bnr_Hy;
  const bool bnr = bnr_y != nullptr;

  alignas(64) CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)(q.rows_per_tile + 2), 1};
    if (int rc = tc::make_tmap_bf16(&ta, x, 4, dims, strides, box, "conv activation halo map")) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout};
    const uint64_t strides[1] = {(uint64_t)9 * Cin * 2};
    const uint32_t box[2] = {64, (uint32_t)q.kN};
    if (int rc = tc::make_tmap_bf16(&tb, wpack, 2, dims, strides, box, "conv weight map")) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
#define SED_HALO(N, MT, RES) (bnr ? launch_halo<N, MT, RES, true>(ta, tb, p, q, s) : launch_halo<N, MT, RES, false>(ta, tb, p, q, s))
  if (q.kN == 64 && q.resident) return SED_HALO(64, 2, true);
  if (q.kN == 64) return SED_HALO(64, 2, false);
  if (q.kN == 128) return SED_HALO(128, 2, false);
  return SED_HALO(256, 1, false);
#undef SED_HALO
}

int sed_conv3x3_tc_fwd(const void* x, const void* wpack, void* y, float* stats_partial, int B, int H, int W,
                       int Cin, int Cout, sed_stream_t stream) {
  return conv_halo_dispatch(x, wpack, y, stats_partial, B, H, W, Cin, Cout, nullptr, nullptr, nullptr, 0, 0, stream,
                            "sed_conv3x3_tc_fwd");
}

int sed_conv3x3_tc_dgrad_bnr(const void* dy, const void* wpack_dgrad, void* dx, int B, int H, int W, int Cin, int Cout,
                             const void* y_below, int Hy, const float* scale_below, const float* shift_below,
                             int pool, float* partial, sed_stream_t stream) {
  SED_REQUIRE(y_below && scale_below && shift_below && partial, "sed_conv3x3_tc_dgrad_bnr: null pointer");
  SED_REQUIRE(pool == 1 || pool == 2, "sed_conv3x3_tc_dgrad_bnr: pool=%d must be 1 or 2", pool);
  SED_REQUIRE(Hy / pool == H, "sed_conv3x3_tc_dgrad_bnr: y has %d rows per clip, expected %d..%d", Hy, H * pool,
              H * pool + pool - 1);
  SED_REQUIRE(aligned(y_below, 16) && aligned(scale_below, 16) && aligned(shift_below, 16),
              "sed_conv3x3_tc_dgrad_bnr: y / scale / shift must be 16-byte aligned");
  return conv_halo_dispatch(dy, wpack_dgrad, dx, partial, B, H, W, Cin, Cout, y_below, scale_below, shift_below, pool,
                            Hy, stream,
                            "sed_conv3x3_tc_dgrad_bnr");
}

}  // extern "C"
