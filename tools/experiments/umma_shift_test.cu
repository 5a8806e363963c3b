// Experiment (run under gpurun): does a tcgen05 shared-memory matrix descriptor whose start address is
// shifted by whole 128-byte rows (not 1024-byte aligned) read a 128B-swizzled tile consistently with the
// absolute-address swizzle that TMA / manual stores used?  Tested for
//   (1) K-major A  : rows = M index (pixels), 64 bf16 of K per row   -> start += s*128 B  <=> M rows shifted by s
//   (2) MN-major B : rows = K index (pixels), 64 bf16 of N per row   -> start += s*128 B  <=> K rows shifted by s
// each with descriptor base_offset = 0 and base_offset = (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_test umma_shift_test.cu && ./umma_shift_test
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// byte offset of element (row, col) in a [rows][64 bf16] tile with 128B swizzle (absolute-address pattern,
// tile base 1024B aligned)
__host__ __device__ inline int sw_off(int row, int col) {
  const int chunk = (col * 2) / 16, within = (col * 2) % 16;
  return row * 128 + ((chunk ^ (row & 7)) * 16) + within;
}

// mode 0: K-major A shifted (A tile 160 rows x 64 K; B [64 N][64 K] K-major unshifted)
//         D[m][n] = sum_k A[m+s][k] * B[n][k],  M=128, N=64, K=64
// mode 1: MN-major B shifted (B tile 96 rows(K) x 64 N; A [64 K rows][2 slabs of 64 M] MN-major unshifted)
//         D[m][n] = sum_k A[k][m] * B[k+s][n],  M=128, N=64, K=64
__global__ void test_kernel(const __nv_bfloat16* A, const __nv_bfloat16* Bm, float* D, int mode, int shift, int use_base_off) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // up to 160 rows * 128 B = 20 KB (two slabs for mode 1: 2 * 64 rows * 128 = 16 KB)
  uint8_t* sB = smem + 32768;         // up to 96 rows * 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x;
  if (mode == 0) {
    for (int i = tid; i < 160 * 64; i += blockDim.x) { int r = i / 64, c = i % 64; *(__nv_bfloat16*)(sA + sw_off(r, c)) = A[i]; }
    for (int i = tid; i < 64 * 64; i += blockDim.x) { int r = i / 64, c = i % 64; *(__nv_bfloat16*)(sB + sw_off(r, c)) = Bm[i]; }
  } else {
    // A: global [64 k][128 m] -> two MN-major slabs [64 k rows][64 m], slab stride 8192 B
    for (int i = tid; i < 64 * 128; i += blockDim.x) { int k = i / 128, m = i % 128; *(__nv_bfloat16*)(sA + (m / 64) * 8192 + sw_off(k, m % 64)) = A[i]; }
    // B: global [96 k][64 n] -> one slab
    for (int i = tid; i < 96 * 64; i += blockDim.x) { int k = i / 64, n = i % 64; *(__nv_bfloat16*)(sB + sw_off(k, n)) = Bm[i]; }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> async proxy (MMA)
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t da, db;
      uint32_t idesc;
      if (mode == 0) {
        const uint32_t a_addr = smem_u32(sA) + shift * 128 + ks * 32;
        const uint32_t b_addr = smem_u32(sB) + ks * 32;
        da = desc_sw128(a_addr, 16, 1024, use_base_off ? ((a_addr >> 7) & 7) : 0);
        db = desc_sw128(b_addr, 16, 1024, 0);
        idesc = idesc_bf16(128, 64, 0, 0);
      } else {
        const uint32_t a_addr = smem_u32(sA) + ks * 2048;                  // 16 k rows per MMA
        const uint32_t b_addr = smem_u32(sB) + shift * 128 + ks * 2048;
        da = desc_sw128(a_addr, 8192, 1024, 0);
        db = desc_sw128(b_addr, 8192, 1024, use_base_off ? ((b_addr >> 7) & 7) : 0);
        idesc = idesc_bf16(128, 64, 1, 1);
      }
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0; int spins = 0;
    while (!ok && spins < (1 << 24)) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      ++spins;
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5, lane = tid & 31;     // 4 warps: lanes 32*warp..
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 32) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * 64 + c * 32 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<float> hA(160 * 64 > 64 * 128 ? 160 * 64 : 64 * 128), hB(96 * 64);
  srand(1);
  for (auto& v : hA) v = bf((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : hB) v = bf((rand() % 2001 - 1000) / 1000.0f);
  std::vector<__nv_bfloat16> bA(hA.size()), bB(hB.size());
  for (size_t i = 0; i < hA.size(); ++i) bA[i] = __float2bfloat16(hA[i]);
  for (size_t i = 0; i < hB.size(); ++i) bB[i] = __float2bfloat16(hB[i]);
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, bA.size() * 2); cudaMalloc(&dB, bB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hD(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 17, 31};
  for (int mode = 0; mode < 2; ++mode)
    for (int ubo = 0; ubo < 2; ++ubo)
      for (int s : shifts) {
        cudaMemset(dD, 0, 128 * 64 * 4);
        test_kernel<<<1, 128, 64 * 1024>>>(dA, dB, dD, mode, s, ubo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d base_off %d shift %d: CUDA error %s\n", mode, ubo, s, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k)
              ref += mode == 0 ? (double)hA[(m + s) * 64 + k] * hB[n * 64 + k] : (double)hA[k * 128 + m] * hB[(k + s) * 64 + n];
            maxerr = fmax(maxerr, fabs(ref - hD[m * 64 + n]));
          }
        printf("mode %d (%s) base_off=%s shift %2d rows: max abs err %.3e  %s\n", mode, mode == 0 ? "K-major A shifted" : "MN-major B shifted",
               ubo ? "(addr>>7)&7" : "0", s, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
      }
  return 0;
}
