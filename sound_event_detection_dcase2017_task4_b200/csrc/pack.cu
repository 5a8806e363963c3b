// pack.cu -- layout / precision shadows of the reference's fp32 master tensors.
#include "common.cuh"

namespace sed {
namespace {

// OIHW fp32 (Cout, Cin, 3, 3)  ->  fwd   [Cout][tap][Cin]            bf16   (k = tap*Cin + ci)
//                              ->  dgrad [Cin][8 - tap][Cout]        bf16   (180-degree rotated, transposed)
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, int Cout, int Cin,
                                         __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgrad) {
  const long long n = (long long)Cout * Cin * 9;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % Cin);
    const int co = (int)(i / (9LL * Cin));
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (fwd) fwd[((long long)co * 9 + tap) * Cin + ci] = v;
    if (dgrad) dgrad[((long long)ci * 9 + (8 - tap)) * Cout + co] = v;
  }
}

// wgrad result [tap][Cout][Cin] fp32 (one or more split-K slabs summed) -> OIHW fp32 gradient
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ g, int slabs, long long slab_stride, int Cout,
                                         int Cin, float* __restrict__ out, int accumulate) {
  const long long n = (long long)Cout * Cin * 9;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % Cin);
    const int co = (int)(i / (9LL * Cin));
    const long long src = ((long long)tap * Cout + co) * Cin + ci;
    float acc = 0.f;
    for (int s = 0; s < slabs; ++s) acc += g[s * slab_stride + src];
    out[i] = accumulate ? out[i] + acc : acc;
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_conv_pack_weights(const float* w_oihw, int Cout, int Cin, void* fwd_pack, void* dgrad_pack,
                          sed_stream_t stream) {
  SED_REQUIRE(w_oihw && (fwd_pack || dgrad_pack), "sed_conv_pack_weights: null pointer");
  const long long n = (long long)Cout * Cin * 9;
  const int grid = (int)min((long long)sm_count() * 8, (n + 255) / 256);
  pack_conv_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(fwd_pack), reinterpret_cast<__nv_bfloat16*>(dgrad_pack));
  SED_LAUNCH_CHECK("pack_conv_weights_kernel");
  return 0;
}

int sed_conv_unpack_wgrad(const float* g_tap_major, int slabs, long long slab_stride, int Cout, int Cin,
                          float* grad_oihw, int accumulate, sed_stream_t stream) {
  SED_REQUIRE(g_tap_major && grad_oihw && slabs >= 1, "sed_conv_unpack_wgrad: bad arguments");
  const long long n = (long long)Cout * Cin * 9;
  const int grid = (int)min((long long)sm_count() * 8, (n + 255) / 256);
  unpack_conv_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g_tap_major, slabs, slab_stride, Cout, Cin,
                                                                   grad_oihw, accumulate);
  SED_LAUNCH_CHECK("unpack_conv_wgrad_kernel");
  return 0;
}

int sed_f32_to_bf16(const float* x, void* y, long long n, sed_stream_t stream) {
  SED_REQUIRE(x && y && n >= 0, "sed_f32_to_bf16: bad arguments");
  if (n == 0) return 0;
  const int grid = (int)min((long long)sm_count() * 8, (n + 255) / 256);
  f32_to_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n);
  SED_LAUNCH_CHECK("f32_to_bf16_kernel");
  return 0;
}

}  // extern "C"
