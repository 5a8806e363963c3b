// common.cuh -- shared host/device helpers for libsedb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/sed_b200.h"

namespace sed {

// ---- error plumbing (no exceptions across the C ABI) --------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SED_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::sed::set_error(__VA_ARGS__);      \
      return 1;                           \
    }                                     \
  } while (0)

#define SED_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::sed::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                       \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

// call right after a <<<>>> launch
#define SED_LAUNCH_CHECK(name)                                                        \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::sed::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));      \
      return 3;                                                                       \
    }                                                                                 \
    ::sed::count_launch();                                                            \
  } while (0)

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
int sm_count();   // of the current device (cached per device)

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Deterministic second-level reduction used by every "finalize" kernel: block = (32 columns, kColLanes row lanes).
// Row lane y sums rows y, y + kColLanes, ... of two [P][stride] fp32 partial arrays (a, b) for column `col`
// in fp64; the lanes are then combined in fixed order.  The totals are valid in the threadIdx.y == 0 threads.
constexpr int kColLanes = 16;
__device__ __forceinline__ void colsum2_block(const float* __restrict__ a, const float* __restrict__ b, int P,
                                              long long stride, long long col, bool ok, double& sa, double& sb) {
  __shared__ double s_part[2][kColLanes][33];
  double xa = 0.0, xb = 0.0;
  if (ok) {
    // four rows per trip with the loads issued before the (sequential, fixed-order) fp64 adds: the loop is a chain of
    // L2 round trips otherwise (37 of them for the 592 partial rows of a convolution: 17 us per finalize)
    int p = threadIdx.y;
    for (; p + 3 * kColLanes < P; p += 4 * kColLanes) {
      float va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        va[u] = a[(long long)(p + u * kColLanes) * stride + col];
        vb[u] = b ? b[(long long)(p + u * kColLanes) * stride + col] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xa += (double)va[u];
        xb += (double)vb[u];
      }
    }
    for (; p < P; p += kColLanes) {
      xa += (double)a[(long long)p * stride + col];
      if (b) xb += (double)b[(long long)p * stride + col];
    }
  }
  s_part[0][threadIdx.y][threadIdx.x] = xa;
  s_part[1][threadIdx.y][threadIdx.x] = xb;
  __syncthreads();
  sa = 0.0; sb = 0.0;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int l = 0; l < kColLanes; ++l) {
      sa += s_part[0][l][threadIdx.x];
      sb += s_part[1][l][threadIdx.x];
    }
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) {
  return __uint_as_float(((uint32_t)b) << 16);
}
#endif

}  // namespace sed
