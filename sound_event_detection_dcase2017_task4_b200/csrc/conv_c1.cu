// conv_c1.cu -- the first convolution of the network, Cin = 1 (ConvBlock1.conv1,
// /root/reference/pytorch/models.py:181 ctor, :102 forward): K = 9 per output -- an HBM-bound
// stencil, not a GEMM -- so it runs on the CUDA cores with the 9 x 8 weights of a thread's channel
// octet held in registers.  Input fp32 (B,H,W) (the bn0/SpecAug/mixup output), output NHWC bf16
// (B,H,W,Cout) + the per-channel BatchNorm statistics of the fp32 result.
// Backward: weight gradient (Cout,1,3,3) and the input gradient (needed only for bn0's affine
// parameters).
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

// dW[co][tap] partials: same mapping, 8 x 9 accumulators per thread.
__global__ void __launch_bounds__(kThreadsC1)
conv_c1_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int B, int H, int W, int Cout,
                     float* __restrict__ partial /* [grid][Cout*9] */) {
  extern __shared__ float smem[];
  const int ldx = W + 2;
  float* sx = smem;                                  // [(kRows+2)][W+2]
  float* s_red = smem + (kRows + 2) * ldx;           // [lanes][Cout*9]
  const int OV = Cout / 8;
  const int lanes = kThreadsC1 / OV;
  const int ov = threadIdx.x % OV, pl = threadIdx.x / OV;
  float acc[8][9];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
  const int chunks = (H + kRows - 1) / kRows;
  const int items = B * chunks;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / chunks, h0 = (item - b * chunks) * kRows;
    const int nr = min(kRows, H - h0);
    __syncthreads();
    stage_rows(x + (long long)b * H * W, H, W, h0 - 1, nr + 2, sx);
    __syncthreads();
#pragma unroll 2
    for (int idx = pl; idx < nr * W; idx += lanes) {
      const int r = idx / W, wq = idx - r * W;
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(dy + (((long long)b * H + h0 + r) * W + wq) * Cout + ov * 8));
      const float* c = sx + r * ldx + wq;
      float in[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) in[kh * 3 + kw] = c[kh * ldx + kw];
      float g[8];
      float2 q0 = unpack_bf16x2(raw.x), q1 = unpack_bf16x2(raw.y), q2 = unpack_bf16x2(raw.z), q3 = unpack_bf16x2(raw.w);
      g[0] = q0.x; g[1] = q0.y; g[2] = q1.x; g[3] = q1.y; g[4] = q2.x; g[5] = q2.y; g[6] = q3.x; g[7] = q3.y;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[k][t] = fmaf(g[k], in[t], acc[k][t]);
    }
  }
  const int n = Cout * 9;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) s_red[pl * n + (ov * 8 + k) * 9 + t] = acc[k][t];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * n + i];
    partial[(long long)blockIdx.x * n + i] = a;
  }
}

// dX[h][w] = sum_{kh,kw} T[kh*3+kw][h-kh+1][w-kw+1],  T[tap][p] = sum_co dY[p][co] * w[co][tap].
// Per work item the CTA computes T for the chunk's rows plus one halo row either side (each dY row is
// read once per item: 1.25x in total).  One thread per pixel: it streams the pixel's Cout bf16 values
// (16-byte loads, the whole warp covers a contiguous 32 x Cout x 2 B span) against the 9 x Cout weight
// table broadcast from shared memory -- no cross-thread reduction -- then the 9 shifted planes are
// summed from shared memory.
constexpr int kRowsD = 8;
__global__ void __launch_bounds__(kThreadsC1, 2)
conv_c1_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w, int B, int H, int W, int Cout,
                     float* __restrict__ dx) {
  extern __shared__ __align__(16) float smem_d[];
  const int ldx = W + 2, plane = (kRowsD + 2) * ldx;
  float* sT = smem_d;                                // [9][(kRowsD+2)][W+2], zero border
  float* sWt = smem_d + ((9 * plane + 3) & ~3);      // [Cout/8][9][8]: w[co][tap] regrouped by channel octet
  for (int i = threadIdx.x; i < Cout * 9; i += kThreadsC1) {
    const int co = i / 9, t = i - co * 9;
    sWt[((co >> 3) * 9 + t) * 8 + (co & 7)] = w[i];
  }
  const int OV = Cout / 8;
  const int chunks = (H + kRowsD - 1) / kRowsD;
  const int items = B * chunks;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / chunks, h0 = (item - b * chunks) * kRowsD;
    const int nr = min(kRowsD, H - h0);
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * plane; i += kThreadsC1) sT[i] = 0.f;
    __syncthreads();
    // T for image rows h0-1 .. h0+nr (halo row index r = h - (h0-1))
    for (int idx = threadIdx.x; idx < (nr + 2) * W; idx += kThreadsC1) {
      const int r = idx / W, wq = idx - r * W;
      const int h = h0 - 1 + r;
      if (h < 0 || h >= H) continue;
      const uint4* src = reinterpret_cast<const uint4*>(dy + (((long long)b * H + h) * W + wq) * Cout);
      float t[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) t[i] = 0.f;
#pragma unroll 4
      for (int o = 0; o < OV; ++o) {
        const uint4 raw = __ldg(src + o);
        float g[8];
        float2 q0 = unpack_bf16x2(raw.x), q1 = unpack_bf16x2(raw.y), q2 = unpack_bf16x2(raw.z), q3 = unpack_bf16x2(raw.w);
        g[0] = q0.x; g[1] = q0.y; g[2] = q1.x; g[3] = q1.y; g[4] = q2.x; g[5] = q2.y; g[6] = q3.x; g[7] = q3.y;
        const float4* wp = reinterpret_cast<const float4*>(sWt + o * 72);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const float4 wa = wp[2 * i], wb = wp[2 * i + 1];
          t[i] = fmaf(g[0], wa.x, t[i]); t[i] = fmaf(g[1], wa.y, t[i]); t[i] = fmaf(g[2], wa.z, t[i]); t[i] = fmaf(g[3], wa.w, t[i]);
          t[i] = fmaf(g[4], wb.x, t[i]); t[i] = fmaf(g[5], wb.y, t[i]); t[i] = fmaf(g[6], wb.z, t[i]); t[i] = fmaf(g[7], wb.w, t[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) sT[i * plane + r * ldx + wq + 1] = t[i];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nr * W; idx += kThreadsC1) {
      const int r = idx / W, wq = idx - r * W;           // output row h0 + r  <-> halo row r + 1
      float a = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
          // y[h'][w'] used x[h'+kh-1][w'+kw-1]  =>  dx[h][w] += T[kh,kw][h-kh+1][w-kw+1]
          a += sT[(kh * 3 + kw) * plane + (r + 1 - kh + 1) * ldx + (wq - kw + 1) + 1];
      dx[((long long)b * H + h0 + r) * W + wq] = a;
    }
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
  const int lanes = kThreadsC1 / (Cout / 8);
  const size_t smem_fwd = ((size_t)(kRows + 2) * (W + 2) + (size_t)lanes * 2 * Cout) * sizeof(float);
  SED_CUDA(cudaFuncSetAttribute(conv_c1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SED_REQUIRE(smem_fwd <= 160 * 1024, "sed_conv_c1_fwd: shared memory");
  conv_c1_fwd_kernel<<<sed_conv_c1_grid(), kThreadsC1, smem_fwd, (cudaStream_t)stream>>>(
      x, w, B, H, W, Cout, reinterpret_cast<__nv_bfloat16*>(y), stats_partial);
  SED_LAUNCH_CHECK("conv_c1_fwd_kernel");
  return 0;
}

int sed_conv_c1_wgrad(const float* x, const void* dy, float* partial, int B, int H, int W, int Cout,
                      sed_stream_t stream) {
  SED_REQUIRE(x && dy && partial, "sed_conv_c1_wgrad: null pointer");
  SED_REQUIRE(Cout % 8 == 0 && Cout <= 128 && 256 % (Cout / 8) == 0, "sed_conv_c1_wgrad: Cout=%d unsupported", Cout);
  const int lanes = kThreadsC1 / (Cout / 8);
  SED_REQUIRE(W >= 1 && W <= 1024, "sed_conv_c1_wgrad: W=%d out of range", W);
  const size_t smem = ((size_t)(kRows + 2) * (W + 2) + (size_t)lanes * Cout * 9) * sizeof(float);
  SED_CUDA(cudaFuncSetAttribute(conv_c1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SED_REQUIRE(smem <= 160 * 1024, "sed_conv_c1_wgrad: shared memory");
  conv_c1_wgrad_kernel<<<sed_conv_c1_grid(), kThreadsC1, smem, (cudaStream_t)stream>>>(
      x, reinterpret_cast<const __nv_bfloat16*>(dy), B, H, W, Cout, partial);
  SED_LAUNCH_CHECK("conv_c1_wgrad_kernel");
  return 0;
}

int sed_conv_c1_dgrad(const void* dy, const float* w, float* dx, int B, int H, int W, int Cout, sed_stream_t stream) {
  SED_REQUIRE(dy && w && dx, "sed_conv_c1_dgrad: null pointer");
  const int OV = Cout / 8;
  SED_REQUIRE(Cout % 8 == 0 && OV >= 1 && OV <= 32 && (OV & (OV - 1)) == 0, "sed_conv_c1_dgrad: Cout=%d unsupported", Cout);
  if (B == 0) return 0;
  SED_REQUIRE(W >= 1 && W <= 256 && (kThreadsC1 / OV) >= 1, "sed_conv_c1_dgrad: W=%d out of range", W);
  const size_t smem_d = ((((size_t)9 * (kRowsD + 2) * (W + 2) + 3) & ~(size_t)3) + (size_t)Cout * 9) * sizeof(float);
  SED_CUDA(cudaFuncSetAttribute(conv_c1_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  SED_REQUIRE(smem_d <= 160 * 1024, "sed_conv_c1_dgrad: shared memory");
  conv_c1_dgrad_kernel<<<sm_count() * 2, kThreadsC1, smem_d, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), w, B, H, W, Cout, dx);
  SED_LAUNCH_CHECK("conv_c1_dgrad_kernel");
  return 0;
}

}  // extern "C"
