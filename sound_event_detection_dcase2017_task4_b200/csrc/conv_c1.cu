// conv_c1.cu -- the first convolution of the network, Cin = 1 (ConvBlock1.conv1,
// /root/reference/pytorch/models.py:181 ctor, :102 forward): K = 9 per output, an HBM-bound stencil.
// Input fp32 (B,H,W) (the bn0/SpecAug/mixup output), output NHWC bf16 (B,H,W,Cout) + the per-channel
// BatchNorm statistics of the fp32 result.  Two forms: the reference configuration (Cout = 64, W % 16 == 0) runs as
// warp-level tensor-core tiles (conv_c1_fwd_mma_kernel, below); any other shape on the CUDA cores with the 9 x 8
// weights of a thread's channel octet in registers (conv_c1_fwd_kernel).
// Backward: weight gradient (Cout,1,3,3) here; the input gradient (needed only for bn0's affine
// parameters) is a tcgen05 kernel in conv_c1_tc.cu.
#include "common.cuh"

namespace sed {
namespace {

constexpr int kThreadsC1 = 256;

// Common structure of the three kernels: a CTA walks (clip, chunk of kRows image rows) work items;
// the fp32 input rows of the chunk (+1 halo row either side, +1 zero column either side) are staged
// in shared memory, so the 3x3 window reads are conflict-free LDS and the zero padding costs no
// branches.  thread = (pixel lane, channel octet); octets vary fastest so the 8 threads of a pixel
// read / write one contiguous Cout*2-byte NHWC row (16-byte vectors).
constexpr int kRows = 8;

__device__ __forceinline__ void stage_rows(const float* __restrict__ img, int H, int W, int h0, int nrows,
                                           float* __restrict__ sx) {
  const int ldx = W + 2;
  for (int i = threadIdx.x; i < nrows * ldx; i += kThreadsC1) {
    const int r = i / ldx, c = i - r * ldx;
    const int h = h0 + r, w = c - 1;
    sx[i] = (h >= 0 && h < H && w >= 0 && w < W) ? __ldg(img + (long long)h * W + w) : 0.f;
  }
}

__global__ void __launch_bounds__(kThreadsC1)
conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w /* (Cout, 9) */, int B, int H, int W,
                   int Cout, __nv_bfloat16* __restrict__ y, float* __restrict__ partial) {
  extern __shared__ float smem[];
  const int ldx = W + 2;
  float* sx = smem;                                  // [(kRows+2)][W+2]
  float* s_red = smem + (kRows + 2) * ldx;           // [lanes][2*Cout]
  const int OV = Cout / 8;
  const int lanes = kThreadsC1 / OV;
  const int ov = threadIdx.x % OV, pl = threadIdx.x / OV;
  float wt[8][9];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[k][t] = w[(ov * 8 + k) * 9 + t];
  float s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = ss[k] = 0.f;
  const int chunks = (H + kRows - 1) / kRows;
  const int items = B * chunks;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / chunks, h0 = (item - b * chunks) * kRows;
    const int nr = min(kRows, H - h0);
    __syncthreads();
    stage_rows(x + (long long)b * H * W, H, W, h0 - 1, nr + 2, sx);
    __syncthreads();
    for (int idx = pl; idx < nr * W; idx += lanes) {
      const int r = idx / W, wq = idx - r * W;
      const float* c = sx + r * ldx + wq;            // window top-left (halo row r <-> image row h0+r-1)
      float in[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) in[kh * 3 + kw] = c[kh * ldx + kw];
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(in[t], wt[k][t], a);
        o[k] = a;
        s[k] += a;
        ss[k] += a * a;
      }
      uint4 pk;
      pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
      pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(y + (((long long)b * H + h0 + r) * W + wq) * Cout + ov * 8) = pk;
    }
  }
  if (partial == nullptr) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s_red[pl * 2 * Cout + ov * 8 + k] = s[k];
    s_red[pl * 2 * Cout + Cout + ov * 8 + k] = ss[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * 2 * Cout + i];
    partial[(long long)blockIdx.x * 2 * Cout + i] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core form (Cout = 64, W a multiple of 16): the stencil as a K = 32 GEMM on warp-level mma.sync tiles.
//   M = 16 consecutive pixels of an image row, N = 64 channels, K = [x_hi(9 taps) | x_lo(9) | x_hi(9) | 0(5)] against
//   [w_hi(9); w_hi(9); w_lo(9); 0]: the three bf16 products whose sum is the fp32 product up to 2^-16 (x and w are fp32
//   in the reference; the lo*lo term is dropped) -- two m16n8k16 MMAs per 8 channels.
// The CUDA-core kernel above issues 9 FMAs + 2 statistics FMAs per output and was bound by fp32 issue (0.95 ms per
// step, 2.6 TB/s); here a warp spends 16 MMAs + a 16-element operand gather per 1024 outputs, and the kernel is left
// with what it must do: stream 2.1 GB of bf16 outputs (floor 0.33 ms).
//   * input rows of a work item (8 image rows + halo) are split ONCE into bf16 hi / lo planes in shared memory;
//   * the weight fragments (8 channel tiles x 2 k-steps) live in registers for the whole kernel;
//   * the fp32 accumulators feed the BatchNorm statistics (packed fp32x2 adds / FMAs) before they are rounded;
//   * the 16 x 64 bf16 tile goes through a per-warp staging buffer so that every lane stores 16 contiguous bytes
//     (4 pixels x 128 B per instruction, fully coalesced NHWC rows).
constexpr int kC1Ld = 72;                 // bf16 elements per staged input row (W + 2 <= 72), and per staged output pixel
constexpr int kC1Plane = (kRows + 2) * kC1Ld;

__device__ __forceinline__ void mma16816_c1(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
// element k (0..31) of the K axis: which plane (0 hi, 1 lo, 2 none) and which tap
__device__ __forceinline__ void c1_k_source(int k, int& plane, int& tap) {
  if (k < 9) { plane = 0; tap = k; }
  else if (k < 18) { plane = 1; tap = k - 9; }
  else if (k < 27) { plane = 0; tap = k - 18; }
  else { plane = 2; tap = 0; }
}
__device__ __forceinline__ float c1_weight_k(const float* __restrict__ w, int n, int k) {
  if (k >= 27) return 0.f;
  const float v = w[n * 9 + (k % 9)];
  const float hi = __bfloat162float(__float2bfloat16_rn(v));
  return k < 18 ? hi : v - hi;                       // rows 0..17: w_hi, rows 18..26: w_lo
}

__global__ void __launch_bounds__(kThreadsC1, 2)
conv_c1_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ w /* (64, 9) */, int B, int H, int W,
                       __nv_bfloat16* __restrict__ y, float* __restrict__ partial) {
  constexpr int Cout = 64;
  __shared__ __align__(16) __nv_bfloat16 s_in[2 * 2 * kC1Plane];          // two buffers of (hi plane, lo plane)
  __shared__ __align__(16) __nv_bfloat16 s_out[(kThreadsC1 / 32) * 16 * kC1Ld];
  __shared__ float s_red[(kThreadsC1 / 32) * 2 * Cout];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  // ---- per-thread constants: weight fragments and the operand gather table
  uint32_t bf[8][2][2];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int sI = 0; sI < 2; ++sI)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int k = 16 * sI + 2 * t4 + 8 * hh, n = 8 * j + g;
        bf[j][sI][hh] = pack_bf16x2(c1_weight_k(w, n, k), c1_weight_k(w, n, k + 1));
      }
  int off[8];                       // element offsets (plane + tap) of k = 16 s + 2 t + 8 hh + e, index ((s*2 + hh)*2 + e)
  uint32_t pair_mask[4];            // 0xFFFF per live half of the packed pair ((s*2 + hh))
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t m = 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = 16 * (q >> 1) + 2 * t4 + 8 * (q & 1) + e;
      int plane, tap;
      c1_k_source(k, plane, tap);
      off[q * 2 + e] = (plane == 1 ? kC1Plane : 0) + (tap / 3) * kC1Ld + (tap % 3);
      if (plane != 2) m |= 0xFFFFu << (16 * e);
    }
    pair_mask[q] = m;
  }
  float2 sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = make_float2(0.f, 0.f);
  __nv_bfloat16* st = s_out + warp * 16 * kC1Ld;
  const int tiles_w = W >> 4;
  const int chunks = (H + kRows - 1) / kRows;
  const int items = B * chunks;
  const int ldx = W + 2;
  // The input rows of item n+1 are fetched into registers while item n computes and written (split into bf16 hi / lo
  // planes, zero borders) into the other shared-memory buffer afterwards: the global-load latency of the tiny input
  // never sits between two barriers.  (kRows + 2) * (W + 2) <= 3 * 256 elements = 3 per thread.
  float pf[3];
  auto fetch = [&](int it) {
    const int b = it / chunks, h0 = (it - b * chunks) * kRows;
    const int nr = min(kRows, H - h0);
    const float* img = x + (long long)b * H * W;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = threadIdx.x + q * kThreadsC1;
      const int r = i / ldx, c = i - r * ldx;
      const int h = h0 - 1 + r, wc = c - 1;
      pf[q] = (r < nr + 2 && h >= 0 && h < H && wc >= 0 && wc < W) ? __ldg(img + (long long)h * W + wc) : 0.f;
    }
  };
  auto stash = [&](int buf) {
    __nv_bfloat16* dst = s_in + buf * 2 * kC1Plane;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = threadIdx.x + q * kThreadsC1;
      const int r = i / ldx, c = i - r * ldx;
      if (r < kRows + 2) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(pf[q]);
        dst[r * kC1Ld + c] = hi;
        dst[kC1Plane + r * kC1Ld + c] = __float2bfloat16_rn(pf[q] - __bfloat162float(hi));
      }
    }
  };
  int cur = 0;
  if ((int)blockIdx.x < items) {
    fetch(blockIdx.x);
    stash(0);
  }
  __syncthreads();
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / chunks, h0 = (item - b * chunks) * kRows;
    const int nr = min(kRows, H - h0);
    const int next = item + gridDim.x;
    if (next < items) fetch(next);
    const unsigned short* s_in16 = reinterpret_cast<const unsigned short*>(s_in + cur * 2 * kC1Plane);
    for (int tile = warp; tile < nr * tiles_w; tile += kThreadsC1 / 32) {
      const int r = tile / tiles_w, w0 = (tile - r * tiles_w) << 4;
      const int base0 = r * kC1Ld + w0 + g, base1 = base0 + 8;       // window top-left of pixels w0+g and w0+g+8
      uint32_t a[2][4];
#pragma unroll
      for (int sI = 0; sI < 2; ++sI)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int q = sI * 2 + hh;
          const uint32_t lo0 = s_in16[base0 + off[q * 2]], hi0 = s_in16[base0 + off[q * 2 + 1]];
          const uint32_t lo1 = s_in16[base1 + off[q * 2]], hi1 = s_in16[base1 + off[q * 2 + 1]];
          a[sI][2 * hh] = (lo0 | (hi0 << 16)) & pair_mask[q];          // row g,     k = 16 s + 2 t + 8 hh (+1)
          a[sI][2 * hh + 1] = (lo1 | (hi1 << 16)) & pair_mask[q];      // row g + 8
        }
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        mma16816_c1(acc[j], a[0], bf[j][0][0], bf[j][0][1]);
        mma16816_c1(acc[j], a[1], bf[j][1][0], bf[j][1][1]);
      }
      __syncwarp();                                                    // previous tile's staging reads are done
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 top = make_float2(acc[j][0], acc[j][1]), bot = make_float2(acc[j][2], acc[j][3]);
        sum[j] = add2(sum[j], add2(top, bot));
        sq[j] = fma2(top, top, fma2(bot, bot, sq[j]));
        *reinterpret_cast<uint32_t*>(st + g * kC1Ld + 8 * j + 2 * t4) = pack_bf16x2(top.x, top.y);
        *reinterpret_cast<uint32_t*>(st + (g + 8) * kC1Ld + 8 * j + 2 * t4) = pack_bf16x2(bot.x, bot.y);
      }
      __syncwarp();
      __nv_bfloat16* dst = y + (((long long)b * H + h0 + r) * W + w0) * Cout;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int px = i * 4 + (lane >> 3), ch = (lane & 7) * 8;
        *reinterpret_cast<uint4*>(dst + px * Cout + ch) = *reinterpret_cast<const uint4*>(st + px * kC1Ld + ch);
      }
    }
    if (next < items) stash(cur ^ 1);
    __syncthreads();
    cur ^= 1;
  }
  if (partial == nullptr) return;
  // ---- per-CTA statistics row: lanes with the same t4 hold the same channels -> fold the 8 row groups, then the warps
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      sum[j].x += __shfl_xor_sync(0xffffffffu, sum[j].x, o);
      sum[j].y += __shfl_xor_sync(0xffffffffu, sum[j].y, o);
      sq[j].x += __shfl_xor_sync(0xffffffffu, sq[j].x, o);
      sq[j].y += __shfl_xor_sync(0xffffffffu, sq[j].y, o);
    }
    if (g == 0) {
      float* row = s_red + warp * 2 * Cout;
      row[8 * j + 2 * t4] = sum[j].x;
      row[8 * j + 2 * t4 + 1] = sum[j].y;
      row[Cout + 8 * j + 2 * t4] = sq[j].x;
      row[Cout + 8 * j + 2 * t4 + 1] = sq[j].y;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cout; i += kThreadsC1) {
    float v = 0.f;
    for (int wv = 0; wv < kThreadsC1 / 32; ++wv) v += s_red[wv * 2 * Cout + i];
    partial[(long long)blockIdx.x * 2 * Cout + i] = v;
  }
}

}  // namespace
}  // namespace sed


using namespace sed;

extern "C" {

int sed_conv_c1_grid(void) { return sm_count() * 4; }

int sed_conv_c1_fwd(const float* x, const float* w, void* y, float* stats_partial, int B, int H, int W, int Cout,
                    sed_stream_t stream) {
  SED_REQUIRE(x && w && y, "sed_conv_c1_fwd: null pointer");
  SED_REQUIRE(Cout % 8 == 0 && Cout <= 256 && 256 % (Cout / 8) == 0, "sed_conv_c1_fwd: Cout=%d unsupported", Cout);
  if (B == 0) return 0;
  SED_REQUIRE(W >= 1 && W <= 1024, "sed_conv_c1_fwd: W=%d out of range", W);
  if (Cout == 64 && W % 16 == 0 && W + 2 <= kC1Ld && (kRows + 2) * (W + 2) <= 3 * kThreadsC1 && aligned(y, 16)) {
    conv_c1_fwd_mma_kernel<<<sed_conv_c1_grid(), kThreadsC1, 0, (cudaStream_t)stream>>>(
        x, w, B, H, W, reinterpret_cast<__nv_bfloat16*>(y), stats_partial);
    SED_LAUNCH_CHECK("conv_c1_fwd_mma_kernel");
    return 0;
  }
  const int lanes = kThreadsC1 / (Cout / 8);
  const size_t smem_fwd = ((size_t)(kRows + 2) * (W + 2) + (size_t)lanes * 2 * Cout) * sizeof(float);
  SED_CUDA(cudaFuncSetAttribute(conv_c1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SED_REQUIRE(smem_fwd <= 160 * 1024, "sed_conv_c1_fwd: shared memory");
  conv_c1_fwd_kernel<<<sed_conv_c1_grid(), kThreadsC1, smem_fwd, (cudaStream_t)stream>>>(
      x, w, B, H, W, Cout, reinterpret_cast<__nv_bfloat16*>(y), stats_partial);
  SED_LAUNCH_CHECK("conv_c1_fwd_kernel");
  return 0;
}

}  // extern "C"
