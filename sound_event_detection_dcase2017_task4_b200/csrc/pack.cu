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

// ---- all layers of a model in ONE launch (the per-layer entry points above cost 7 + 7 tiny launches per step) ----
constexpr int kMaxPackLayers = 8;
struct PackTable {
  const float* w[kMaxPackLayers];          // OIHW fp32 weights
  __nv_bfloat16* fwd[kMaxPackLayers];      // may be null
  __nv_bfloat16* dgrad[kMaxPackLayers];    // may be null
  int cout[kMaxPackLayers], cin[kMaxPackLayers];
  int first_block[kMaxPackLayers + 1];     // blocks [first_block[l], first_block[l + 1]) work on layer l
  int layers;
};
__global__ void pack_conv_weights_multi_kernel(const __grid_constant__ PackTable t) {
  int l = 0;
  while (l + 1 < t.layers && (int)blockIdx.x >= t.first_block[l + 1]) ++l;
  const int Cout = t.cout[l], Cin = t.cin[l];
  const long long n = (long long)Cout * Cin * 9;
  const float* __restrict__ w = t.w[l];
  __nv_bfloat16* __restrict__ fwd = t.fwd[l];
  __nv_bfloat16* __restrict__ dgrad = t.dgrad[l];
  const int nb = t.first_block[l + 1] - t.first_block[l];
  for (long long i = ((long long)blockIdx.x - t.first_block[l]) * blockDim.x + threadIdx.x; i < n;
       i += (long long)nb * blockDim.x) {
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % Cin);
    const int co = (int)(i / (9LL * Cin));
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (fwd) fwd[((long long)co * 9 + tap) * Cin + ci] = v;
    if (dgrad) dgrad[((long long)ci * 9 + (8 - tap)) * Cout + co] = v;
  }
}

struct UnpackTable {
  const float* slabs[kMaxPackLayers];      // [n_slabs][tap][Cout][Cin] fp32
  float* out[kMaxPackLayers];              // OIHW fp32 gradient
  int n_slabs[kMaxPackLayers], cout[kMaxPackLayers], cin[kMaxPackLayers];
  int first_block[kMaxPackLayers + 1];
  int layers;
};
// reads run along Cin (coalesced: every slab is read once, there are up to 16 of them), the 4-byte writes are strided
__global__ void unpack_conv_wgrad_multi_kernel(const __grid_constant__ UnpackTable t) {
  int l = 0;
  while (l + 1 < t.layers && (int)blockIdx.x >= t.first_block[l + 1]) ++l;
  const int Cout = t.cout[l], Cin = t.cin[l], S = t.n_slabs[l];
  const long long n = (long long)Cout * Cin * 9;
  const float* __restrict__ g = t.slabs[l];
  float* __restrict__ out = t.out[l];
  const int nb = t.first_block[l + 1] - t.first_block[l];
  for (long long j = ((long long)blockIdx.x - t.first_block[l]) * blockDim.x + threadIdx.x; j < n;
       j += (long long)nb * blockDim.x) {
    const int ci = (int)(j % Cin);
    const int co = (int)((j / Cin) % Cout);
    const int tap = (int)(j / ((long long)Cin * Cout));
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += g[s * n + j];        // same slab order as unpack_conv_wgrad_kernel: bit-identical
    out[((long long)co * Cin + ci) * 9 + tap] = acc;
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

int sed_conv_pack_weights_multi(int layers, const float* const* w_oihw, const int* Cout, const int* Cin, void* const* fwd_pack,
                                void* const* dgrad_pack, sed_stream_t stream) {
  SED_REQUIRE(layers >= 1 && layers <= kMaxPackLayers && w_oihw && Cout && Cin && fwd_pack && dgrad_pack,
              "sed_conv_pack_weights_multi: bad arguments (1..%d layers)", kMaxPackLayers);
  PackTable t;
  t.layers = layers;
  int blocks = 0;
  for (int l = 0; l < layers; ++l) {
    SED_REQUIRE(w_oihw[l] && (fwd_pack[l] || dgrad_pack[l]), "sed_conv_pack_weights_multi: null pointer (layer %d)", l);
    t.w[l] = w_oihw[l];
    t.fwd[l] = reinterpret_cast<__nv_bfloat16*>(fwd_pack[l]);
    t.dgrad[l] = reinterpret_cast<__nv_bfloat16*>(dgrad_pack[l]);
    t.cout[l] = Cout[l]; t.cin[l] = Cin[l];
    t.first_block[l] = blocks;
    const long long n = (long long)Cout[l] * Cin[l] * 9;
    blocks += (int)min((long long)sm_count() * 2, (n + 1023) / 1024);
  }
  t.first_block[layers] = blocks;
  pack_conv_weights_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t);
  SED_LAUNCH_CHECK("pack_conv_weights_multi_kernel");
  return 0;
}

int sed_conv_unpack_wgrad_multi(int layers, const float* const* slabs, const int* n_slabs, const int* Cout, const int* Cin,
                                float* const* grad_oihw, sed_stream_t stream) {
  SED_REQUIRE(layers >= 1 && layers <= kMaxPackLayers && slabs && n_slabs && Cout && Cin && grad_oihw,
              "sed_conv_unpack_wgrad_multi: bad arguments (1..%d layers)", kMaxPackLayers);
  UnpackTable t;
  t.layers = layers;
  int blocks = 0;
  for (int l = 0; l < layers; ++l) {
    SED_REQUIRE(slabs[l] && grad_oihw[l] && n_slabs[l] >= 1, "sed_conv_unpack_wgrad_multi: bad layer %d", l);
    t.slabs[l] = slabs[l]; t.out[l] = grad_oihw[l];
    t.n_slabs[l] = n_slabs[l]; t.cout[l] = Cout[l]; t.cin[l] = Cin[l];
    t.first_block[l] = blocks;
    const long long n = (long long)Cout[l] * Cin[l] * 9;
    blocks += (int)min((long long)sm_count() * 2, (n + 1023) / 1024);
  }
  t.first_block[layers] = blocks;
  unpack_conv_wgrad_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t);
  SED_LAUNCH_CHECK("unpack_conv_wgrad_multi_kernel");
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
