// conv_c1_tc.cu -- backward of the first convolution (Cin = 1, Cout = 64; ConvBlock1.conv1,
// /root/reference/pytorch/models.py:181 ctor, :102 forward; autograd backward at pytorch/main.py:257) on tcgen05.
//
// Data gradient (needed for bn0's affine parameters):
//   dX[h][w] = sum_{kh,kw} T[kh*3+kw][h-kh+1][w-kw+1],     T[tap][p] = sum_co dY[p][co] * w[co][tap]
// T is a (pixels x 64) x (64 x 9) GEMM: the 2.1 GB dY tensor streams through TMA exactly like the A operand of the
// 3x3 kernels (K-major SW128 tiles, OOB rows zero-filled), the 9 taps are the N dimension.  The fp32 weights are
// split into bf16 hi + lo parts stacked in N (columns 0-8 and 16-24 of a 32-column accumulator), so the product
// keeps fp32-class accuracy (dY is bf16 already).  The CUDA-core version this replaces spent 1.44 ms L1-bound
// (89 % L1/TEX: 576 FMAs + 18 weight LDS.128 per pixel and uncoalesced 16-byte reads); here the kernel is bound by
// the one pass over dY.
//
// CTA tile = R output rows + 1 halo row either side (kMT M tiles of 128 pixels); warp 0 = TMA producer, warp 1 = MMA
// issuer, warps 2-9 = epilogue: tcgen05.ld -> T planes in shared memory (zero borders = the conv padding) ->
// 9-term shifted sum -> fp32 dX.  Accumulators double buffered in TMEM.
#include "common.cuh"
#include "tc.cuh"

namespace sed {
namespace {

using namespace tc;

constexpr int kC1Threads = 320;
constexpr int kC1Epi = 256;
constexpr int kC1MT = 5;                       // M tiles (128 pixels each) per CTA tile
constexpr int kC1N = 32;                       // accumulator columns: [hi taps 0..8 | pad | lo taps 16..24 | pad]
constexpr int kC1Stages = 2;

struct C1Params {
  int B, H, W;
  int bh;            // image rows per M tile (128 / W)
  int rows_tile;     // kC1MT * bh  (= R + 2)
  int R;             // output rows per CTA tile
  int tiles_h, num_tiles;
  int a_bytes;       // rows_tile * W * 128
  const float* w;    // (64, 9)
  float* dx;         // (B, H, W)
};

// packed fp32 pair FMA (one issue slot for two IEEE fmas: the fused apply transform is issue-bound)
__device__ __forceinline__ float2 c1_fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ int sw128_off(int row, int col_bf16) {
  const int chunk = (col_bf16 * 2) >> 4, within = (col_bf16 * 2) & 15;
  return row * 128 + ((chunk ^ (row & 7)) << 4) + within;
}

__global__ void __launch_bounds__(kC1Threads, 1)
conv_c1_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const C1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                          // kC1Stages x a_bytes
  uint8_t* smem_b = smem + kC1Stages * p.a_bytes;                  // [32][64] bf16 K-major SW128 (4 KB)
  float* sT = reinterpret_cast<float*>(smem_b + 4096);             // [9][rows_tile][W + 2]
  __shared__ uint64_t a_full[kC1Stages], a_empty[kC1Stages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  constexpr uint32_t kTmemCols = 512;                              // 2 x kC1MT x 32 = 320 -> next power of two

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldx = p.W + 2, plane = p.rows_tile * ldx;

  // weights -> B operand (rows = accumulator column n, 64 K-elements = output channels), bf16 hi / lo split
  for (int i = threadIdx.x; i < kC1N * 64; i += kC1Threads) {
    const int n = i >> 6, co = i & 63;
    const int tap = n & 15;
    float v = 0.f;
    if (tap < 9) {
      const float wv = p.w[co * 9 + tap];
      const float hi = __bfloat162float(__float2bfloat16_rn(wv));
      v = n < 16 ? hi : wv - hi;
    }
    *reinterpret_cast<__nv_bfloat16*>(smem_b + sw128_off(n, co)) = __float2bfloat16_rn(v);
  }
  for (int i = threadIdx.x; i < 9 * plane; i += kC1Threads) sT[i] = 0.f;
  fence_proxy_async();                                             // generic-proxy smem writes -> visible to UMMA
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_dy);
    for (int s = 0; s < kC1Stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], kC1Epi); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int b = tile / p.tiles_h, h0 = (tile % p.tiles_h) * p.R;
        mbar_wait(&a_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&a_full[st], (uint32_t)p.a_bytes);
        tma_load_4d(smem_a + st * p.a_bytes, &tmap_dy, &a_full[st], 0, 0, h0 - 1, b);
        if (++st == kC1Stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kC1N, 0, 0);
      const uint32_t sb = smem_u32(smem_b);
      int st = 0, it = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait(&a_full[st], ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem_a + st * p.a_bytes);
#pragma unroll
        for (int mt = 0; mt < kC1MT; ++mt) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = umma_desc_sw128(sa + mt * 16384 + ks * 32, 16, 1024);
            const uint64_t db = umma_desc_sw128(sb + ks * 32, 16, 1024);
            umma_bf16(tmem_base + (acc * kC1MT + mt) * kC1N, da, db, idesc, ks != 0 ? 1u : 0u);
          }
        }
        umma_commit(&a_empty[st]);
        umma_commit(&tmem_full_bar[acc]);
        if (++st == kC1Stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;                               // 0..255 within the epilogue group
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int b = tile / p.tiles_h, h0 = (tile % p.tiles_h) * p.R;
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      for (int mt = half; mt < kC1MT; mt += 2) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (acc * kC1MT + mt) * kC1N, v);
        const int pix = mt * 128 + q * 32 + lane;                  // pixel of the (rows_tile x W) tile
        const int r = pix / p.W, wq = pix - r * p.W;
        float* dst = sT + r * ldx + wq + 1;
#pragma unroll
        for (int t = 0; t < 9; ++t) dst[t * plane] = v[t] + v[16 + t];
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      asm volatile("bar.sync 1, %0;" ::"n"(kC1Epi) : "memory");    // all T planes of this tile are written
      const int nr = min(p.R, p.H - h0);
      for (int idx = et; idx < nr * p.W; idx += kC1Epi) {
        const int r = idx / p.W, wq = idx - r * p.W;               // output row h0 + r  <-> tile row r + 1
        float a = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            // y[h'][w'] used x[h'+kh-1][w'+kw-1]  =>  dx[h][w] += T[kh,kw][h-kh+1][w-kw+1]
            a += sT[(kh * 3 + kw) * plane + (r + 1 - kh + 1) * ldx + (wq - kw + 1) + 1];
        p.dx[((long long)b * p.H + h0 + r) * p.W + wq] = a;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kC1Epi) : "memory");    // planes free for the next tile
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// Weight gradient:  dW[co][tap] = sum_p dY[p][co] * x0[p + tap offset]   (zero padding)
// A GEMM with K = all B*H*W pixels, M = 64 output channels, N = 9 taps:
//   * A = dY, NHWC bf16: a TMA box {64 ch, W, bh, 1} IS a canonical MN-major SW128 slab (rows = pixels = K index),
//     the same operand form as the 3x3 weight-gradient kernels; M = 128 is issued with LBO = 0, so accumulator rows
//     64..127 duplicate rows 0..63 and are ignored;
//   * B = im2col(x0) is built in shared memory by the eight worker warps from fp32 input rows staged with zero
//     borders: per pixel the 9 window values are split into bf16 hi + lo parts (columns 0-8 and 16-24), one 128 B
//     MN-major row, 128B-swizzled with ordinary 16-byte stores;
//   * every CTA accumulates its whole pixel range into ONE 128 x 32 fp32 TMEM accumulator and writes one partial
//     [64][9] block at the end (summed in fixed order by sed_reduce_partials).
// The CUDA-core version spent 1.12 ms (8 x 9 FMAs per pixel-octet, latency-bound at 25 % occupancy); this one is
// bound by the single pass over dY.
constexpr int kWgStages = 4;
constexpr int kWgRows = 16;           // image rows per work item (x0 rows staged once per item)
constexpr int kWgSxFloats = (kWgRows + 2) * (128 + 2);   // one sx buffer (W <= 128)
constexpr int kWgPf = ((kWgRows + 2) * (128 + 2) + kC1Epi - 1) / kC1Epi;   // prefetch registers per builder thread (W = 128 worst case)

struct C1WgParams {
  int B, H, W, bh;
  int chunks, items;                  // row chunks per clip, B * chunks
  const float* x0;                    // (B, H, W) fp32
  float* partial;                     // [gridDim.x][64 * 9]
  // fused BatchNorm-backward apply (kFused): dY is made here from y and dA and also written out for the data gradient
  const float* scale; const float* shift; const float* mean; const float* invstd; const float* coef;   // (64,) ... (3, 64)
  __nv_bfloat16* dy_out;              // (B, H, W, 64)
};

// kFused: the A slabs arriving by TMA are the raw conv output y (tmap_dy) and the incoming gradient dA (tmap_g) of
// block1.bn1; the builder group that makes a block's im2col tile first turns the two slabs IN PLACE into the dY slab
// (sed_bn_relu_pool_bwd_apply's arithmetic: element i of both slabs is the same (pixel, channel) whatever the swizzle, and
// a thread's 16-byte chunks all hold the same 8 channels, whose constants it keeps in registers) and stores it to global
// memory for the data-gradient kernel.  One pass over y and dA replaces the apply pass plus this kernel's own pass over dY.
template <bool kFused>
__global__ void __launch_bounds__(kC1Threads, 1)
conv_c1_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_g,
                        const C1WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                          // kWgStages x 16 KB   dY slabs (TMA)
  uint8_t* smem_b = smem + kWgStages * 16384;                      // kWgStages x 16 KB   im2col tiles (built here)
  uint8_t* smem_g = smem_b + kWgStages * 16384;                    // kFused: kWgStages x 16 KB   dA slabs (TMA)
  float* sx = reinterpret_cast<float*>(smem_g + (kFused ? kWgStages * 16384 : 0));   // [kWgRows + 2][W + 2] fp32 input rows
  __shared__ uint64_t a_full[kWgStages], b_full[kWgStages], empty_bar[kWgStages], done_bar;
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldx = p.W + 2;
  const int kb_per_item_max = kWgRows / p.bh;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_dy);
    if (kFused) tma_prefetch_desc(&tmap_g);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&b_full[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<32>(&tmem_base_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int b = item / p.chunks, h0 = (item % p.chunks) * kWgRows;
        const int nkb = (min(kWgRows, p.H - h0) + p.bh - 1) / p.bh;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[st], ph ^ 1);
          mbar_arrive_expect_tx(&a_full[st], kFused ? 32768u : 16384u);
          tma_load_4d(smem_a + st * 16384, &tmap_dy, &a_full[st], 0, 0, h0 + kb * p.bh, b);   // rows >= H: zeros
          if (kFused) tma_load_4d(smem_g + st * 16384, &tmap_g, &a_full[st], 0, 0, h0 + kb * p.bh, b);
          if (++st == kWgStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kC1N, 1, 1);
      int st = 0;
      uint32_t ph = 0;
      bool first = true;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int h0 = (item % p.chunks) * kWgRows;
        const int nkb = (min(kWgRows, p.H - h0) + p.bh - 1) / p.bh;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&a_full[st], ph);
          mbar_wait(&b_full[st], ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem_a + st * 16384), sb = smem_u32(smem_b + st * 16384);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                         // 128 pixels = 8 x K16
            const uint64_t da = umma_desc_sw128(sa + ks * 2048, 0, 1024);
            const uint64_t db = umma_desc_sw128(sb + ks * 2048, 0, 1024);
            umma_bf16(tmem_base, da, db, idesc, (first && ks == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit(&empty_bar[st]);
          if (++st == kWgStages) { st = 0; ph ^= 1; }
        }
      }
      umma_commit(&done_bar);
    }
  } else {
    const int et = threadIdx.x - 64;                               // 0..255
    const int grp = et >> 7, gt = et & 127;                        // builder group, thread (= pixel) within it
    int nblk = 0;                                                  // K blocks seen so far (all items of this CTA)
    bool any = false;
    // The fp32 input rows of item n+1 are fetched into registers while the tiles of item n are built, and written to
    // the OTHER sx buffer at the top of the next iteration: one barrier per item and no global-load latency between
    // barriers (round 1 staged them synchronously: the builders -- and behind them the MMA and the TMA ring -- idled
    // ~1 us per 1.5 us item, which is why the kernel sat at 0.5 of the HBM roofline).
    float pf[kWgPf];
    auto fetch = [&](int it) {
      const int b = it / p.chunks, h0 = (it % p.chunks) * kWgRows;
      const float* img = p.x0 + (long long)b * p.H * p.W;
#pragma unroll
      for (int q = 0; q < kWgPf; ++q) {
        const int i = et + q * kC1Epi;
        const int r = i / ldx, c = i - r * ldx;
        const int h = h0 - 1 + r, w = c - 1;
        pf[q] = (r < kWgRows + 2 && h >= 0 && h < p.H && w >= 0 && w < p.W) ? __ldg(img + (long long)h * p.W + w) : 0.f;
      }
    };
    // kFused: the 8 channels of this thread's 16-byte chunks (logical chunk = physical chunk ^ (row & 7); a thread's
    // chunks are 16 rows apart, so the logical chunk is the same for all of them) and their apply constants
    const int lc = (gt & 7) ^ ((gt >> 3) & 7);
    float2 k_sc[4], k_sh[4], k_a[4], k_b[4], k_c[4];               // channel pairs (2 e2, 2 e2 + 1) of the chunk
    if (kFused) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = lc * 8 + e;
        const float a = p.coef[c], c2 = p.coef[64 + c], c3 = p.coef[128 + c];
        const float kb = -a * c3 * p.invstd[c];
        const float kc = -a * c2 - kb * p.mean[c];
        if (e & 1) { k_sc[e >> 1].y = p.scale[c]; k_sh[e >> 1].y = p.shift[c]; k_a[e >> 1].y = a; k_b[e >> 1].y = kb; k_c[e >> 1].y = kc; }
        else { k_sc[e >> 1].x = p.scale[c]; k_sh[e >> 1].x = p.shift[c]; k_a[e >> 1].x = a; k_b[e >> 1].x = kb; k_c[e >> 1].x = kc; }
      }
    }
    int cur = 0;
    if ((int)blockIdx.x < p.items) fetch(blockIdx.x);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int h0 = (item % p.chunks) * kWgRows;
      const int nr = min(kWgRows, p.H - h0);
      const int nkb = (nr + p.bh - 1) / p.bh;
      float* sxc = sx + cur * kWgSxFloats;
#pragma unroll
      for (int q = 0; q < kWgPf; ++q) {
        const int i = et + q * kC1Epi;
        if (i < (kWgRows + 2) * ldx) sxc[i] = pf[q];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kC1Epi) : "memory");    // sx[cur] is complete; sx[cur ^ 1] has no reader left
      if (item + (int)gridDim.x < p.items) fetch(item + gridDim.x);
      cur ^= 1;
      for (int kb = 0; kb < nkb; ++kb, ++nblk) {
        // the two 4-warp groups build alternate K blocks, so two tiles are under construction at any time
        if ((nblk & 1) == grp) {
          const int st = nblk % kWgStages;
          const uint32_t ph = (uint32_t)(nblk / kWgStages) & 1u;
          mbar_wait(&empty_bar[st], ph ^ 1);
          if (kFused) {
            mbar_wait(&a_full[st], ph);                            // y and dA slabs of this block have landed
            uint4* ys = reinterpret_cast<uint4*>(smem_a + st * 16384);
            const uint4* gs = reinterpret_cast<const uint4*>(smem_g + st * 16384);
            const int bimg = item / p.chunks;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int i = gt + 128 * k;                          // chunk index: slab row i >> 3 (pixel), chunk i & 7
              const int prow = i >> 3;
              const int hh = h0 + kb * p.bh + prow / p.W, ww = prow % p.W;
              const uint4 yv = ys[i], gv = gs[i];
              const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
              uint32_t ow[4];
#pragma unroll
              for (int e2 = 0; e2 < 4; ++e2) {
                const float2 xf = unpack_bf16x2(yw[e2]), gf = unpack_bf16x2(gw[e2]);
                const float2 act = c1_fma2(xf, k_sc[e2], k_sh[e2]);                  // bn(y): the ReLU mask
                const float2 gm = make_float2(act.x > 0.f ? gf.x : 0.f, act.y > 0.f ? gf.y : 0.f);
                const float2 o = c1_fma2(k_a[e2], gm, c1_fma2(k_b[e2], xf, k_c[e2]));
                ow[e2] = pack_bf16x2(o.x, o.y);
              }
              const bool ok = hh < p.H;                            // rows beyond the image: dY = 0 (what TMA's zero fill gave)
              const uint4 ov = ok ? make_uint4(ow[0], ow[1], ow[2], ow[3]) : make_uint4(0u, 0u, 0u, 0u);
              ys[i] = ov;
              if (ok)
                *reinterpret_cast<uint4*>(p.dy_out + (((long long)bimg * p.H + hh) * p.W + ww) * 64 + lc * 8) = ov;
            }
          }
          uint8_t* tile = smem_b + st * 16384;
          const int r = kb * p.bh + gt / p.W, wq = gt % p.W;       // pixel row within the item, column
          const float* c = sxc + r * ldx + wq;                     // window top-left (staged row r <-> image row h0+r-1)
          uint32_t hi[5], lo[5];
          float win[10];
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) win[kh * 3 + kw] = c[kh * ldx + kw];
          win[9] = 0.f;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const float a0 = win[2 * i], a1 = win[2 * i + 1];
            const float h0f = __bfloat162float(__float2bfloat16_rn(a0)), h1f = __bfloat162float(__float2bfloat16_rn(a1));
            hi[i] = pack_bf16x2(h0f, h1f);
            lo[i] = pack_bf16x2(a0 - h0f, a1 - h1f);
          }
          // logical 16-byte chunks of this pixel's MN-major row: 0 = taps 0-7 hi, 1 = tap 8 hi, 2 = taps 0-7 lo, 3 = tap 8 lo
          uint4* row = reinterpret_cast<uint4*>(tile + gt * 128);
          const int sw = gt & 7;
          row[0 ^ sw] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          row[1 ^ sw] = make_uint4(hi[4], 0u, 0u, 0u);
          row[2 ^ sw] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          row[3 ^ sw] = make_uint4(lo[4], 0u, 0u, 0u);
          fence_proxy_async();
          mbar_arrive(&b_full[st]);
        }
        any = true;
      }
    }
    (void)kb_per_item_max;
    // epilogue: accumulator rows 0..63 = output channels (warps with lane quarter 0 and 1 of the first epilogue group)
    if (warp >= 2 && warp < 6 && (warp & 3) < 2) {
      const int q = warp & 3;
      float v[32];
      if (any) {
        mbar_wait(&done_bar, 0);
        tcgen05_fence_after();
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      float* out = p.partial + (long long)blockIdx.x * 576 + (q * 32 + lane) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) out[t] = v[t] + v[16 + t];
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_conv_c1_dgrad(const void* dy, const float* w, float* dx, int B, int H, int W, int Cout, sed_stream_t stream) {
  SED_REQUIRE(dy && w && dx, "sed_conv_c1_dgrad: null pointer");
  SED_REQUIRE(Cout == 64, "sed_conv_c1_dgrad: Cout=%d unsupported (the tensor-core path is built for 64)", Cout);
  SED_REQUIRE(W >= 8 && W <= 128 && 128 % W == 0, "sed_conv_c1_dgrad: W=%d must divide 128 and be >= 8", W);
  if (B == 0) return 0;
  SED_REQUIRE(H >= 1 && (long long)B * H < (1LL << 31), "sed_conv_c1_dgrad: bad shape");
  C1Params p;
  p.B = B; p.H = H; p.W = W;
  p.bh = 128 / W;
  p.rows_tile = kC1MT * p.bh;
  p.R = p.rows_tile - 2;
  SED_REQUIRE(p.R >= 1, "sed_conv_c1_dgrad: W=%d leaves no output rows per tile", W);
  p.tiles_h = (H + p.R - 1) / p.R;
  p.num_tiles = B * p.tiles_h;
  p.a_bytes = p.rows_tile * W * 128;
  p.w = w; p.dx = dx;
  alignas(64) CUtensorMap tm;
  {
    const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)p.rows_tile, 1};
    if (int rc = tc::make_tmap_bf16(&tm, dy, 4, dims, strides, box, "conv_c1 dY map")) return rc;
  }
  const size_t smem = (size_t)kC1Stages * p.a_bytes + 4096 + (size_t)9 * p.rows_tile * (W + 2) * sizeof(float) + 1024;
  SED_REQUIRE(smem <= 220 * 1024, "sed_conv_c1_dgrad: shared memory");
  SED_CUDA(cudaFuncSetAttribute(conv_c1_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  conv_c1_dgrad_tc_kernel<<<grid, kC1Threads, smem, (cudaStream_t)stream>>>(tm, p);
  SED_LAUNCH_CHECK("conv_c1_dgrad_tc_kernel");
  return 0;
}

namespace {
int launch_c1_wgrad(const char* name, const float* x, const void* a_src, const void* g_src, C1WgParams p, int B, int H, int W,
                    cudaStream_t stream) {
  SED_REQUIRE(W >= 16 && W <= 128 && 128 % W == 0, "%s: W=%d must divide 128 and be >= 16", name, W);
  SED_REQUIRE(B >= 1 && H >= 1 && (long long)B * H < (1LL << 31), "%s: bad shape", name);
  p.B = B; p.H = H; p.W = W; p.bh = 128 / W;
  p.chunks = (H + kWgRows - 1) / kWgRows;
  p.items = B * p.chunks;
  p.x0 = x;
  const bool fused = g_src != nullptr;
  alignas(64) CUtensorMap tm, tg;
  const uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
  const uint32_t box[4] = {64, (uint32_t)W, (uint32_t)p.bh, 1};
  if (int rc = tc::make_tmap_bf16(&tm, a_src, 4, dims, strides, box, "conv_c1 wgrad dY / y map")) return rc;
  if (int rc = tc::make_tmap_bf16(&tg, fused ? g_src : a_src, 4, dims, strides, box, "conv_c1 wgrad dA map")) return rc;
  const size_t smem = (size_t)(fused ? 3 : 2) * kWgStages * 16384 + (size_t)2 * kWgSxFloats * sizeof(float) + 1024;
  SED_REQUIRE(smem <= 227 * 1024, "%s: shared memory", name);
  // every one of the sed_conv_c1_grid() partial rows is written (CTAs without work write zeros)
  if (fused) {
    SED_CUDA(cudaFuncSetAttribute(conv_c1_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_c1_wgrad_tc_kernel<true><<<sed_conv_c1_grid(), kC1Threads, smem, stream>>>(tm, tg, p);
  } else {
    SED_CUDA(cudaFuncSetAttribute(conv_c1_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_c1_wgrad_tc_kernel<false><<<sed_conv_c1_grid(), kC1Threads, smem, stream>>>(tm, tg, p);
  }
  SED_LAUNCH_CHECK("conv_c1_wgrad_tc_kernel");
  return 0;
}
}  // namespace

int sed_conv_c1_wgrad(const float* x, const void* dy, float* partial, int B, int H, int W, int Cout,
                      sed_stream_t stream) {
  SED_REQUIRE(x && dy && partial, "sed_conv_c1_wgrad: null pointer");
  SED_REQUIRE(Cout == 64, "sed_conv_c1_wgrad: Cout=%d unsupported (the tensor-core path is built for 64)", Cout);
  C1WgParams p{};
  p.partial = partial;
  return launch_c1_wgrad("sed_conv_c1_wgrad", x, dy, nullptr, p, B, H, W, (cudaStream_t)stream);
}

int sed_bn_apply_conv_c1_wgrad(const float* x, const void* y, const void* dA, const float* scale, const float* shift,
                               const float* mean, const float* invstd, const float* coef, void* dy, float* partial, int B,
                               int H, int W, int Cout, sed_stream_t stream) {
  SED_REQUIRE(x && y && dA && scale && shift && mean && invstd && coef && dy && partial,
              "sed_bn_apply_conv_c1_wgrad: null pointer");
  SED_REQUIRE(Cout == 64, "sed_bn_apply_conv_c1_wgrad: Cout=%d unsupported (the tensor-core path is built for 64)", Cout);
  SED_REQUIRE(aligned(dy, 16), "sed_bn_apply_conv_c1_wgrad: dy must be 16-byte aligned");
  C1WgParams p{};
  p.partial = partial;
  p.scale = scale; p.shift = shift; p.mean = mean; p.invstd = invstd; p.coef = coef;
  p.dy_out = reinterpret_cast<__nv_bfloat16*>(dy);
  return launch_c1_wgrad("sed_bn_apply_conv_c1_wgrad", x, y, dA, p, B, H, W, (cudaStream_t)stream);
}

}  // extern "C"
