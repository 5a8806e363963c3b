// conv_c1.cu -- the first convolution of the network, Cin = 1 (ConvBlock1.conv1,
// /root/reference/pytorch/models.py:181 ctor, :102 forward): K = 9 per output -- an HBM-bound
// stencil, not a GEMM -- so it runs on the CUDA cores with the 9 x 8 weights of a thread's channel
// octet held in registers.  Input fp32 (B,H,W) (the bn0/SpecAug/mixup output), output NHWC bf16
// (B,H,W,Cout) + the per-channel BatchNorm statistics of the fp32 result.
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

}  // extern "C"
