// philox.cuh -- counter-based Philox4x32-10 random bits for the two dropouts of MultiHead
// (/root/reference/pytorch/models.py:590 attention dropout 0.1, :652 output dropout 0.2).  A mask is a pure function of
// (seed, offset, element index), so the backward recomputes it instead of storing it.  Statistically, not bit-wise,
// equivalent to torch's own Philox usage.
#pragma once
#include <stdint.h>

namespace sed {

__device__ __forceinline__ uint32_t philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi) {
  const unsigned long long p = (unsigned long long)a * b;
  hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
// the four 32-bit words of block `ctr`
__device__ __forceinline__ uint4 philox4x32_10(unsigned long long seed, unsigned long long ctr) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5ed5ed5eu, c3 = 0;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = philox_mulhilo(0xD2511F53u, c0, hi0);
    const uint32_t lo1 = philox_mulhilo(0xCD9E8D57u, c2, hi1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// word (idx & 3) of block (idx >> 2) + offset
__device__ __forceinline__ uint32_t philox_word(unsigned long long seed, unsigned long long offset,
                                                unsigned long long idx) {
  const uint4 w = philox4x32_10(seed, (idx >> 2) + offset);
  const uint32_t k = (uint32_t)(idx & 3);
  return k == 0 ? w.x : (k == 1 ? w.y : (k == 2 ? w.z : w.w));
}
__device__ __forceinline__ float philox_uniform(uint32_t word) { return (float)(word >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ bool keep_elem(unsigned long long seed, unsigned long long offset,
                                          unsigned long long idx, float p) {
  return philox_uniform(philox_word(seed, offset, idx)) >= p;
}

}  // namespace sed
