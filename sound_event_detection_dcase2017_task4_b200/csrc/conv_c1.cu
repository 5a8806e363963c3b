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

// thread = (pixel lane, channel octet); octets vary fastest so the 8 threads of a pixel write one
// contiguous Cout*2-byte row.
__global__ void __launch_bounds__(kThreadsC1)
conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w /* (Cout, 9) */, int B, int H, int W,
                   int Cout, __nv_bfloat16* __restrict__ y, float* __restrict__ partial) {
  extern __shared__ float s_red[];                   // [lanes][2*Cout]
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
  const long long npix = (long long)B * H * W;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  if (pl < lanes)
    for (long long p = p0 + pl; p < p1; p += lanes) {
      const int wq = (int)(p % W);
      const long long r = p / W;
      const int h = (int)(r % H);
      const float* img = x + (r / H) * (long long)H * W;
      float in[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int hh = h + kh - 1, ww = wq + kw - 1;
          in[kh * 3 + kw] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(img + (long long)hh * W + ww) : 0.f;
        }
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
      *reinterpret_cast<uint4*>(y + p * Cout + ov * 8) = pk;
    }
  if (partial == nullptr) return;
  if (pl < lanes) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s_red[pl * 2 * Cout + ov * 8 + k] = s[k];
      s_red[pl * 2 * Cout + Cout + ov * 8 + k] = ss[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * 2 * Cout + i];
    partial[(long long)blockIdx.x * 2 * Cout + i] = a;
  }
}

// dW[co][tap] partials: same thread mapping, 8 x 9 accumulators per thread.
__global__ void __launch_bounds__(kThreadsC1)
conv_c1_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int B, int H, int W, int Cout,
                     float* __restrict__ partial /* [grid][Cout*9] */) {
  extern __shared__ float s_red[];                   // [lanes][Cout*9]
  const int OV = Cout / 8;
  const int lanes = kThreadsC1 / OV;
  const int ov = threadIdx.x % OV, pl = threadIdx.x / OV;
  float acc[8][9];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
  const long long npix = (long long)B * H * W;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  if (pl < lanes)
    for (long long p = p0 + pl; p < p1; p += lanes) {
      const int wq = (int)(p % W);
      const long long r = p / W;
      const int h = (int)(r % H);
      const float* img = x + (r / H) * (long long)H * W;
      float in[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int hh = h + kh - 1, ww = wq + kw - 1;
          in[kh * 3 + kw] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(img + (long long)hh * W + ww) : 0.f;
        }
      const uint4 raw = *reinterpret_cast<const uint4*>(dy + p * Cout + ov * 8);
      float g[8];
      float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
      g[0] = a.x; g[1] = a.y; g[2] = b.x; g[3] = b.y; g[4] = c.x; g[5] = c.y; g[6] = d.x; g[7] = d.y;
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[k][t] = fmaf(g[k], in[t], acc[k][t]);
    }
  const int n = Cout * 9;
  if (pl < lanes) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t) s_red[pl * n + (ov * 8 + k) * 9 + t] = acc[k][t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * n + i];
    partial[(long long)blockIdx.x * n + i] = a;
  }
}

// dX[p] = sum_tap sum_co dY[p - off(tap)][co] * w[co][tap]; 8 threads (octets) per pixel + shuffle reduce.
__global__ void __launch_bounds__(kThreadsC1)
conv_c1_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ w, int B, int H, int W, int Cout,
                     float* __restrict__ dx) {
  const int OV = Cout / 8;                           // power of two <= 32 (checked on the host)
  const int ov = threadIdx.x % OV;
  float wt[8][9];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[k][t] = w[(ov * 8 + k) * 9 + t];
  const long long npix = (long long)B * H * W;
  const long long stride = (long long)gridDim.x * (blockDim.x / OV);
  const long long iters = (npix + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {
    const long long p = it * stride + (long long)blockIdx.x * (blockDim.x / OV) + threadIdx.x / OV;
    float a = 0.f;
    if (p < npix) {
      const int wq = (int)(p % W);
      const long long r = p / W;
      const int h = (int)(r % H);
      const long long base = (r / H) * (long long)H * W;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          // y[h'][w'] used x[h'+kh-1][w'+kw-1]  =>  dx[h][w] += dy[h-kh+1][w-kw+1] * w[kh][kw]
          const int hh = h - kh + 1, ww = wq - kw + 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
            const uint4 raw = *reinterpret_cast<const uint4*>(dy + (base + (long long)hh * W + ww) * Cout + ov * 8);
            float2 q0 = unpack_bf16x2(raw.x), q1 = unpack_bf16x2(raw.y), q2 = unpack_bf16x2(raw.z),
                   q3 = unpack_bf16x2(raw.w);
            const int t = kh * 3 + kw;
            a = fmaf(q0.x, wt[0][t], a); a = fmaf(q0.y, wt[1][t], a);
            a = fmaf(q1.x, wt[2][t], a); a = fmaf(q1.y, wt[3][t], a);
            a = fmaf(q2.x, wt[4][t], a); a = fmaf(q2.y, wt[5][t], a);
            a = fmaf(q3.x, wt[6][t], a); a = fmaf(q3.y, wt[7][t], a);
          }
        }
    }
    for (int o = OV >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (p < npix && ov == 0) dx[p] = a;
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
  const int lanes = kThreadsC1 / (Cout / 8);
  conv_c1_fwd_kernel<<<sed_conv_c1_grid(), kThreadsC1, (size_t)lanes * 2 * Cout * sizeof(float), (cudaStream_t)stream>>>(
      x, w, B, H, W, Cout, reinterpret_cast<__nv_bfloat16*>(y), stats_partial);
  SED_LAUNCH_CHECK("conv_c1_fwd_kernel");
  return 0;
}

int sed_conv_c1_wgrad(const float* x, const void* dy, float* partial, int B, int H, int W, int Cout,
                      sed_stream_t stream) {
  SED_REQUIRE(x && dy && partial, "sed_conv_c1_wgrad: null pointer");
  SED_REQUIRE(Cout % 8 == 0 && Cout <= 128 && 256 % (Cout / 8) == 0, "sed_conv_c1_wgrad: Cout=%d unsupported", Cout);
  const int lanes = kThreadsC1 / (Cout / 8);
  const size_t smem = (size_t)lanes * Cout * 9 * sizeof(float);
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
  conv_c1_dgrad_kernel<<<sm_count() * 8, kThreadsC1, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), w, B, H, W, Cout, dx);
  SED_LAUNCH_CHECK("conv_c1_dgrad_kernel");
  return 0;
}

}  // extern "C"
