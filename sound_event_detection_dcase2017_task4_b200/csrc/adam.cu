// adam.cu -- one fused multi-tensor Adam(amsgrad=True) update over the flat parameter buffer.
//
// Replaces `optimizer.step()` of `optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.,
// amsgrad=True)` (/root/reference/pytorch/main.py:144-145, :258): torch's per-tensor update
//   m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;  vmax = max(vmax, v)
//   p -= (lr / (1 - b1^t)) * m / (sqrt(vmax) / sqrt(1 - b2^t) + eps)
// as ONE HBM-bound pass over the same flat fp32 gradient buffer the NCCL all-reduce uses.
// grad_scale folds the 1/world_size of the data-parallel mean into the read.
#include "common.cuh"

namespace sed {
namespace {

__global__ void adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, float* __restrict__ vmax, long long n, float lr, float b1,
                                    float b2, float eps, float bc1, float bc2_sqrt, float grad_scale,
                                    const float* __restrict__ bias_corr_dev) {
  if (bias_corr_dev) {                       // step-dependent scalars from device memory (CUDA-graph replays)
    bc1 = bias_corr_dev[0];
    bc2_sqrt = bias_corr_dev[1];
  }
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    const float vm = fmaxf(vmax[i], vi);
    m[i] = mi;
    v[i] = vi;
    vmax[i] = vm;
    p[i] -= step_size * (mi / (sqrtf(vm) / bc2_sqrt + eps));
  }
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_adam_amsgrad(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                     long long n, float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                     const float* bias_corr_dev, sed_stream_t stream) {
  SED_REQUIRE(param && grad && exp_avg && exp_avg_sq && max_exp_avg_sq, "sed_adam_amsgrad: null pointer");
  SED_REQUIRE(step >= 1 && n >= 0, "sed_adam_amsgrad: step must be >= 1");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const int grid = (int)min((n + 255) / 256, (long long)sm_count() * 8);
  adam_amsgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n, lr,
                                                              beta1, beta2, eps, (float)bc1, (float)sqrt(bc2),
                                                              grad_scale, bias_corr_dev);
  SED_LAUNCH_CHECK("adam_amsgrad_kernel");
  return 0;
}

}  // extern "C"
