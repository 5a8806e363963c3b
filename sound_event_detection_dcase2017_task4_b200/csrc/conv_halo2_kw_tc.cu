// conv_halo2_kw_tc.cu -- the CTA-pair halo convolution (conv_halo2_tc.cu) for Cout = 64, with the three kw taps STACKED IN N.
//
// Forward and data gradient of the 64-channel-output 3x3 convolutions (/root/reference/pytorch/models.py:102-103 for
// block 1's conv2, its data gradient and the data gradient of block 2's conv1; backward at pytorch/main.py:257).
// With N = 64 an SS-mode tcgen05.mma is bound by the shared-memory fetch of its A operand: 4 KB per K16 step whatever N
// is, and nine taps re-read the same halo tile nine times (ncu: tensor pipe 56 % active, profiles/r02_step_ncu.md).
// Here one instruction computes all three kw taps of a kernel row:
//     D[p, kw * 32 + c] (+)= sum_ci  x[p + (kh - 1) W, ci] * w[kh, kw, ci, co]          N = 192 over the CTA pair
// with A the UNSHIFTED-in-w halo tile (ONE TMA load per tile instead of three kw-shifted ones) and the kh shift an
// address offset of whole image rows as before: 3 instead of 9 instructions per K16 step, A read 3x instead of 9x.
// The kw shift moves to the output side,
//     y[h, w] = D_kw0[h, w - 1] + D_kw1[h, w] + D_kw2[h, w + 1]                          (zero outside [0, W)),
// i.e. between accumulator ROWS, which are TMEM lanes and belong to different threads of the epilogue: the epilogue
// warps combine them with one shuffle up and one shuffle down per output value, and the two values per warp that cross
// a warp boundary inside an image row (W > 32) go through a 2 KB shared-memory mailbox and a 128-thread named barrier.
//
// N ordering: CTA r of the pair holds B rows [kw][c] = weight rows co = 32 r + c, so accumulator columns
// [96 r + 32 kw, 96 r + 32 kw + 32) are tap kw of output channels 32 r .. 32 r + 31.  The whole 9-tap weight set of a
// CTA (36 KB per 64 input channels) stays resident in shared memory.  Roles, barriers and the 2-deep TMEM ring are those
// of conv_halo2_tc.cu (kMT = 1: 2 x 192 columns).
#include "common.cuh"
#include "tc.cuh"

namespace sed {
namespace {

using namespace tc;

constexpr int kKwThreads = 320;
constexpr int kKwEpi = 256;
constexpr int kKwMaxA = 4;
constexpr int kKwN = 192;
constexpr uint32_t kKwTmemCols = 512;

struct KwParams {
  int B, H, W, Cin;
  int bh;              // image rows per 128-pixel M tile (128 / W); a pair tile covers 2 * bh rows
  int tiles_h, num_tiles;
  int kb;              // 64-channel blocks of Cin
  int a_bytes;         // (bh + 2) * W * 128
  int a_stages;
  __nv_bfloat16* y;
  float* stats;        // [gridDim.x * 4][2][64] or nullptr
};

__device__ __forceinline__ void kw_transpose_reduce2b(float (&a)[32], float (&b)[32], int lane) {
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

// 32 lanes x 16 consecutive fp32 columns, no wait (the caller waits once for several loads)
__device__ __forceinline__ void tmem_ld_32x16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier(int half) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kKwThreads, 1)
conv3x3_halo2_kw_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const KwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.a_stages * p.a_bytes;      // [kh][cb][kw][32 rows][128 B]: 12 KB per (kh, cb)
  __shared__ uint64_t a_full[kKwMaxA], a_empty[kKwMaxA];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2], b_res_full;
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_xch[2][2][4][2][16];  // [sub][half][warp quarter][0: lane 31's kw0, 1: lane 0's kw2][col]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kKwMaxA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 2 * kKwEpi); }
    mbar_init(&b_res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<kKwTmemCols>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs' barriers are initialised before any remote signal
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      const int nb = 9 * p.kb;              // 32-row weight boxes: this CTA's 32 output channels of every (tap, cb)
      if (leader) mbar_arrive_expect_tx(&b_res_full, 2u * (uint32_t)(nb * 32 * 128));
      for (int kh = 0; kh < 3; ++kh)
        for (int cb = 0; cb < p.kb; ++cb)
          for (int kw = 0; kw < 3; ++kw)
            tma_load_2d_2cta(smem_b + (((kh * p.kb + cb) * 3 + kw) * 32) * 128, &tmap_b, &b_res_full,
                             (kh * 3 + kw) * p.Cin + cb * 64, (int)rank * 32);
      int as = 0;
      uint32_t aph = 0;
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs) {
        const int b = tile / p.tiles_h;
        const int h0 = (tile % p.tiles_h) * 2 * p.bh + (int)rank * p.bh;
        for (int cb = 0; cb < p.kb; ++cb) {
          mbar_wait(&a_empty[as], aph ^ 1);
          if (leader) mbar_arrive_expect_tx(&a_full[as], 2u * (uint32_t)p.a_bytes);
          tma_load_4d_2cta(smem_a + as * p.a_bytes, &tmap_a, &a_full[as], cb * 64, 0, h0 - 1, b);
          if (++as == p.a_stages) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, kKwN, 0, 0);
      int as = 0;
      uint32_t aph = 0;
      int it = 0;
      mbar_wait(&b_res_full, 0);
      tcgen05_fence_after();
      for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kKwN;
        for (int cb = 0; cb < p.kb; ++cb) {
          mbar_wait(&a_full[as], aph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem_a + as * p.a_bytes);
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t sb = smem_u32(smem_b + (kh * p.kb + cb) * (3 * 32 * 128));
            const uint32_t sam = sa + (uint32_t)(kh * p.W) * 128u;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = umma_desc_sw128(sam + ks * 32, 16, 1024);
              const uint64_t db = umma_desc_sw128(sb + ks * 32, 16, 1024);
              umma_bf16_2cta(d_tmem, da, db, idesc, (cb | kh | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2cta(&a_empty[as]);
          if (++as == p.a_stages) { as = 0; aph ^= 1; }
        }
        umma_commit_2cta(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs) =====================
    const int q = warp & 3;                             // TMEM lane quarter: pixels 32 q .. 32 q + 31 of the tile
    const int half = (warp - 2) >> 2;                   // output channels 32 (rank) + ... : columns 96 half + 32 kw + ..
    const int row = q * 32 + lane;
    const int r_h = row / p.W, r_w = row % p.W;
    const bool has_left = r_w > 0, has_right = r_w < p.W - 1;
    const bool mailbox = p.W > 32;                      // an image row spans more than one warp
    const float lmask = has_left ? 1.f : 0.f, rmask = has_right ? 1.f : 0.f;
    // warp-uniform: does lane 0's left / lane 31's right neighbour exist (then it lives in the adjacent warp)?
    const bool left_in_mailbox = mailbox && ((q * 32) % p.W) != 0;
    const bool right_in_mailbox = mailbox && ((q * 32 + 32) % p.W) != 0;
    float* stat_row = p.stats == nullptr ? nullptr : p.stats + ((long long)blockIdx.x * 4 + q) * 2 * 64;
    float s_acc[32], ss_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s_acc[i] = ss_acc[i] = 0.f;
    int it = 0;
    for (int tile = pair; tile < p.num_tiles; tile += num_pairs, ++it) {
      const int acc = it & 1;
      const int b = tile / p.tiles_h;
      const int h = (tile % p.tiles_h) * 2 * p.bh + (int)rank * p.bh + r_h;
      const bool valid = h < p.H;
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      __nv_bfloat16* yp = p.y + (((long long)b * p.H + h) * p.W + r_w) * 64 + half * 32;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        uint32_t u0[16], u1[16], u2[16];
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kKwN + half * 96 + sub * 16;
        tmem_ld_32x16_nowait(t0, u0);
        tmem_ld_32x16_nowait(t0 + 32, u1);
        tmem_ld_32x16_nowait(t0 + 64, u2);
        tmem_ld_wait();
        // Values that cross a warp boundary inside an image row (W > 32) go through the mailbox: lane 31 publishes its
        // kw0 row (the next warp's lane 0 needs it), lane 0 its kw2 row; after the barrier lane 31 REPLACES its kw0 registers
        // by the previous warp's (nobody in this warp needs lane 31's own), lane 0 its kw2 registers by the next warp's, so
        // that a lane ROTATION delivers every left / right neighbour, the foreign ones included, with no per-value select.
        if (mailbox) {
          if (lane == 31) {
            uint4* d = reinterpret_cast<uint4*>(&s_xch[sub][half][q][0][0]);
#pragma unroll
            for (int g = 0; g < 4; ++g) d[g] = make_uint4(u0[4 * g], u0[4 * g + 1], u0[4 * g + 2], u0[4 * g + 3]);
          }
          if (lane == 0) {
            uint4* d = reinterpret_cast<uint4*>(&s_xch[sub][half][q][1][0]);
#pragma unroll
            for (int g = 0; g < 4; ++g) d[g] = make_uint4(u2[4 * g], u2[4 * g + 1], u2[4 * g + 2], u2[4 * g + 3]);
          }
          epi_barrier(half);
          if (lane == 31 && left_in_mailbox) {
            const uint4* m = reinterpret_cast<const uint4*>(&s_xch[sub][half][(q + 3) & 3][0][0]);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 v = m[g];
              u0[4 * g] = v.x; u0[4 * g + 1] = v.y; u0[4 * g + 2] = v.z; u0[4 * g + 3] = v.w;
            }
          }
          if (lane == 0 && right_in_mailbox) {
            const uint4* m = reinterpret_cast<const uint4*>(&s_xch[sub][half][(q + 1) & 3][1][0]);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 v = m[g];
              u2[4 * g] = v.x; u2[4 * g + 1] = v.y; u2[4 * g + 2] = v.z; u2[4 * g + 3] = v.w;
            }
          }
        }
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {                  // lmask / rmask: 0 where the neighbour is outside the image row
          const float l = __shfl_sync(0xffffffffu, __uint_as_float(u0[i]), (lane + 31) & 31);
          const float r = __shfl_sync(0xffffffffu, __uint_as_float(u2[i]), (lane + 1) & 31);
          o[i] = fmaf(l, lmask, fmaf(r, rmask, __uint_as_float(u1[i])));
        }
        if (valid) {
          uint4* d4 = reinterpret_cast<uint4*>(yp + sub * 16);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 v;
            v.x = pack_bf16x2(o[8 * g + 0], o[8 * g + 1]);
            v.y = pack_bf16x2(o[8 * g + 2], o[8 * g + 3]);
            v.z = pack_bf16x2(o[8 * g + 4], o[8 * g + 5]);
            v.w = pack_bf16x2(o[8 * g + 6], o[8 * g + 7]);
            d4[g] = v;
          }
          if (stat_row != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              s_acc[sub * 16 + i] += o[i];
              ss_acc[sub * 16 + i] = fmaf(o[i], o[i], ss_acc[sub * 16 + i]);
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive_leader(&tmem_empty_bar[acc]);
    }
    if (stat_row != nullptr) {                          // one reduce per warp: 32 channels of this CTA's 64
      kw_transpose_reduce2b(s_acc, ss_acc, lane);
      stat_row[half * 32 + lane] = s_acc[0];
      stat_row[64 + half * 32 + lane] = ss_acc[0];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                       // no CTA of the pair exits while its peer may still signal its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2cta<kKwTmemCols>(tmem_base);
  }
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_conv3x3_tc2kw_supported(int W, int Cin, int Cout) {
  return (Cout == 64 && Cin % 64 == 0 && Cin >= 64 && Cin <= 128 && W >= 8 && W <= 128 && 128 % W == 0) ? 1 : 0;
}

int sed_conv3x3_tc2kw_fwd(const void* x, const void* wpack, void* y, float* stats_partial, int B, int H, int W, int Cin,
                          int Cout, sed_stream_t stream) {
  SED_REQUIRE(x && wpack && y, "sed_conv3x3_tc2kw_fwd: null pointer");
  SED_REQUIRE(B > 0 && H > 0, "sed_conv3x3_tc2kw_fwd: empty batch");
  SED_REQUIRE(sed_conv3x3_tc2kw_supported(W, Cin, Cout), "sed_conv3x3_tc2kw_fwd: W=%d Cin=%d Cout=%d unsupported", W, Cin,
              Cout);
  SED_REQUIRE(aligned(y, 16), "sed_conv3x3_tc2kw_fwd: output must be 16-byte aligned");
  KwParams p;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin;
  p.bh = 128 / W;
  p.kb = Cin / 64;
  p.a_bytes = (p.bh + 2) * W * 128;
  const int b_bytes = 9 * p.kb * 32 * 128;
  p.a_stages = (216 * 1024 - b_bytes) / p.a_bytes;
  if (p.a_stages > kKwMaxA) p.a_stages = kKwMaxA;
  SED_REQUIRE(p.a_stages >= 2, "sed_conv3x3_tc2kw_fwd: no tiling for W=%d Cin=%d", W, Cin);
  const int smem = b_bytes + p.a_stages * p.a_bytes + 1024;
  p.tiles_h = (H + 2 * p.bh - 1) / (2 * p.bh);
  const long long tiles = (long long)B * p.tiles_h;
  SED_REQUIRE(tiles < (1LL << 31), "sed_conv3x3_tc2kw_fwd: too many tiles");
  p.num_tiles = (int)tiles;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.stats = stats_partial;

  alignas(64) CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)(p.bh + 2), 1};
    if (int rc = tc::make_tmap_bf16(&ta, x, 4, dims, strides, box, "kw-stacked conv activation halo map")) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout};
    const uint64_t strides[1] = {(uint64_t)9 * Cin * 2};
    const uint32_t box[2] = {64, 32};
    if (int rc = tc::make_tmap_bf16(&tb, wpack, 2, dims, strides, box, "kw-stacked conv weight map")) return rc;
  }
  int grid = sm_count() & ~1;
  if ((long long)grid > 2 * tiles) grid = (int)(2 * tiles);
  // every statistics row of the workspace must be defined: rows of CTAs beyond a shortened grid are zeroed here
  if (stats_partial != nullptr && grid < (sm_count() & ~1))
    SED_CUDA(cudaMemsetAsync(stats_partial, 0, sizeof(float) * (size_t)(sm_count() & ~1) * 4 * 2 * Cout, (cudaStream_t)stream));
  SED_CUDA(cudaFuncSetAttribute(conv3x3_halo2_kw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  conv3x3_halo2_kw_kernel<<<grid, kKwThreads, smem, (cudaStream_t)stream>>>(ta, tb, p);
  SED_LAUNCH_CHECK("conv3x3_halo2_kw_kernel");
  return 0;
}

}  // extern "C"
